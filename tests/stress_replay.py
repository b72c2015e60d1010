"""Long replay campaign against the live oracle (CPU, host emulation of the device sources): fresh seeds over the
four movesets (standard, CTCB, linker, linker with two recoil levels), snodin unbound / assembled at eight
temperatures and the 12-domain rasters (linear, cyclic); state bit-exact, tape fully consumed and energy to 1e-12
after every chunk. Not collected by pytest (needs oracle/_ref):  python tests/stress_replay.py SECONDS
Round 1: 9787 runs of 300-600 moves in 900 s, 0 failures. Round 2: 11047 runs, 7 failures - a regression of the FourBody
evaluation (now tests/test_synthetic_systems.py::test_campaign_cases_*); after the fix 9788 runs, 0 failures; with LDO_STRESS_TRACKERS=1 (the Tracked<K> instantiation of the kernels) 6250 runs in
600 s, 0 failures."""
import sys, os, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, ROOT)
import conftest, oracle_ref as o
from latticednaorigami_b200.binding import Simulation
from synthetic import UNIFORM_OPTIONS, write_raster_system
tmp = tempfile.mkdtemp()
t_end = time.time() + float(sys.argv[1])
def run(name, opts, seed, chunks, steps):
    r = o.RefSystem(opts); r.seed(seed)
    sim = Simulation(conftest.write_inp(os.path.join(tmp, f"{seed}.inp"), opts), 1, 0, lib=conftest.load_hostsim())
    if os.environ.get("LDO_STRESS_TRACKERS"): sim.engine.enable_move_trackers(True)  # the Tracked<K> instantiation
    for k in range(chunks):
        r.tape(clear=True); r.simulate(steps); tape = r.tape(clear=True)
        sim.engine.attach_tape(0, tape); sim.engine.run(steps, 0, 0, 0)
        st, info = sim.engine.status()
        if st[0] != 0: return f"status {st[0]} {info[0]} chunk {k}"
        try:
            conftest.assert_state_equal(sim.engine.state(0), r.state(), "x")
            assert sim.engine.tape_position(0) == len(tape), "tape"
            e = r.energy(); assert abs(sim.engine.energies()[0, 0] - e) <= 1e-12 * max(1.0, abs(e)), "energy"
        except AssertionError as ex: return f"MISMATCH {ex} chunk {k}"
    return "ok"
seed = 1000; n = 0; bad = 0
systems = [("snodin_unbound.json", t) for t in (332, 338, 344, 350)] + [("snodin_assembled.json", t) for t in (330, 336, 342, 348)]
movesets = ["moveset_standard.json", "moveset_ctcb.json", "moveset_linker.json", "moveset_linker_heavy.json"]
while time.time() < t_end:
    for ms in movesets:
        for system, temp in systems:
            if time.time() > t_end: break
            seed += 1
            res = run(f"{system} {ms}", conftest.make_options(system, ms, temp=temp), seed, 6, 50)
            n += 1
            if res != "ok": bad += 1; print(system, ms, temp, seed, res, flush=True)
        for cyc in (False, True):
            if time.time() > t_end: break
            seed += 1
            sysf = write_raster_system(os.path.join(tmp, f"r{int(cyc)}.json"), 3, 4, cyc)
            opts = conftest.make_options(temp=302, max_total_staples=8, max_type_staples=2, staple_M=1e-5, **UNIFORM_OPTIONS)
            opts["origami_input_filename"] = sysf; opts["movetype_file"] = os.path.join(conftest.INPUTS, ms)
            res = run(f"raster {cyc} {ms}", opts, seed, 6, 100); n += 1
            if res != "ok": bad += 1; print("raster", cyc, ms, seed, res, flush=True)
print("runs", n, "failures", bad)
