"""A/B throughput of run-kernel build variants (make variant NAME=x DEFS=... -> ab/lib_x.so): the bench
workload at 16384 replicas, 8 warm-up exchange rounds, 4 timed 100-move launches, one subprocess per
library (selected with LDO_B200_LIB). LDO_AB_MODE=assembled times 4144 constant-T replicas started from the
assembled configuration at 330 K instead (the expensive regime). Not a benchmark (see bench.py).
    python profiles/ab.py latticednaorigami_b200/libldo_b200.so ab/lib_x.so ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, tempfile, time
sys.path.insert(0, %r)
import torch
import bench
from latticednaorigami_b200.binding import Simulation
L = len(bench.LADDER)
R = 16384
opts = bench.base_options()
opts.update({"simulation_type": "ut_parallel_tempering", "num_reps": L, "temps": bench.LADDER, "chem_pot_mults": [1] * L,
             "bias_mults": [1] * L, "stacking_mults": [1] * L, "exchange_interval": 100, "swaps": 0, "random_seed": 20261017})
mode = os.environ.get("LDO_AB_MODE", "ptmc")
if mode == "assembled":
    # the expensive regime: constant-T replicas started from the assembled configuration at 330 K
    sys.path.insert(0, os.path.join(%r, "tests"))
    import conftest
    R = 4144
    opts = conftest.make_options("snodin_assembled.json", temp=330, random_seed=7)
    sim = Simulation(bench.write_inp(os.path.join(tempfile.mkdtemp(), "p.inp"), opts), R, 0)
    sim.engine.run(100, 0, 0, 0)
else:
    sim = Simulation(bench.write_inp(os.path.join(tempfile.mkdtemp(), "p.inp"), opts), R, 0)
    for i in range(8):
        sim.engine.run_async(100, 100000, 0, 1000000)
        sim.engine.exchange_collect(to_host=False)
        sim.exchange_apply(i + 1, None)
sim.engine.synchronize()
w0 = time.perf_counter()
for i in range(4):
    sim.engine.run_async(100, 100000, 0, 1000000)
sim.engine.synchronize()
dt = time.perf_counter() - w0
sim.engine.assert_ok()
att, acc = sim.engine.move_stats()
print(f"{mode:9s} {os.path.basename(os.environ.get('LDO_B200_LIB', 'default')):24s} {400 * R / dt / 1e6:8.3f} M moves/s  {dt / 4 * 1e3:8.2f} ms/launch  accepted {acc.sum() / att.sum():.4f}", flush=True)
''' % (ROOT, ROOT)
for lib in sys.argv[1:]:
    env = dict(os.environ, LDO_B200_LIB=os.path.abspath(lib))
    subprocess.run([sys.executable, "-c", CHILD], env=env, check=False)
