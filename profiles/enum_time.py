"""Timing of the device enumeration of examples/enum.inp against the cut depth (prefixes per worker) and the number of workers."""
import os, sys, time, tempfile
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
os.environ["LDO_QUIET"] = "1"
import conftest
from latticednaorigami_b200.binding import Simulation
tmp = tempfile.mkdtemp()
def run(workers, per_worker):
    os.environ["LDO_ENUM_PREFIXES_PER_WORKER"] = str(per_worker)
    o = conftest.make_options("four_unbound.json", temp=340, simulation_type="enumerate", min_total_staples=0, max_total_staples=2, max_type_staples=2,
                              enumerate_staples_only=False, output_filebase=os.path.join(tmp, "e"), ops_to_output="numfulldomains nummisdomains numstackedpairs numstaples")
    t0 = time.time(); sim = Simulation(conftest.write_inp(os.path.join(tmp, "e.inp"), o), workers, 0); t1 = time.time()
    sim.run(); t2 = time.time()
    s = sim.enumeration_summary()
    print(f"workers {workers} prefixes/worker>={per_worker}: create {t1-t0:.2f} s, enumerate {t2-t1:.2f} s, leaves {s['leaves']}, configs {s['num_configs']:.6g}", flush=True)
for workers, pw in [(4144, 8), (4144, 8), (4144, 64), (518, 8), (1036, 8), (2072, 8), (8288, 8)]:
    run(workers, pw)
