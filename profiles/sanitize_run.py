"""Small runs of every kernel for compute-sanitizer (memcheck / racecheck; SURVEY.md 5, race-detection row):
64-replica ensembles of the standard moveset from both snodin starts, the CTCB and linker movesets, a few
replica-exchange rounds and a checkpoint round trip.
    compute-sanitizer --tool racecheck python profiles/sanitize_run.py
"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: E402
from latticednaorigami_b200.binding import Simulation  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
tmp = tempfile.mkdtemp()
cases = [("snodin_unbound.json", "moveset_standard.json", 338, 120), ("snodin_assembled.json", "moveset_standard.json", 330, 40),
         ("snodin_assembled.json", "moveset_ctcb.json", 334, 60), ("snodin_unbound.json", "moveset_linker.json", 338, 80)]
for system, moveset, temp, moves in cases:
    opts = conftest.make_options(system, moveset, temp=temp, random_seed=7)
    sim = Simulation(conftest.write_inp(os.path.join(tmp, "s.inp"), opts), 64, 0)
    sim.engine.run(moves * scale, 50, 0, 60)
    sim.engine.assert_ok()
    sim.engine.energies()
    sim.engine.recompute_energies()
    blob = sim.engine.checkpoint_save()
    sim.engine.checkpoint_load(blob)
    sim.engine.synchronize()
    print(system, moveset, "ok", flush=True)
    sim.close()
opts = conftest.make_options("snodin_unbound.json", simulation_type="ut_parallel_tempering", num_reps=4, temps=[334.0, 337.0, 340.0, 343.0],
                             chem_pot_mults=[1] * 4, bias_mults=[1] * 4, stacking_mults=[1] * 4, exchange_interval=40, swaps=3, random_seed=9)
sim = Simulation(conftest.write_inp(os.path.join(tmp, "pt.inp"), opts), 64, 0)
for swap_i in range(1, 4):
    sim.exchange_round(swap_i)
sim.exchange_state(16, 4)
sim.engine.assert_ok()
print("exchange ok", flush=True)
