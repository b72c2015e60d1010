"""Throughput of the run kernel against the number of RESIDENT warps per SM at a fixed ensemble (16384
replicas, persistent kernel): LDO_MAX_BLOCKS_PER_SM caps the resident blocks (2 warps each). Tells whether
the instruction-fetch stall is latency-like (throughput scales with resident warps) or bandwidth-like
(flat). One subprocess per point because the cap is read at engine creation. Not a benchmark."""
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
R = int(sys.argv[1])
opts = bench.base_options()
opts.update({"simulation_type": "ut_parallel_tempering", "num_reps": L, "temps": bench.LADDER, "chem_pot_mults": [1] * L,
             "bias_mults": [1] * L, "stacking_mults": [1] * L, "exchange_interval": 100, "swaps": 0, "random_seed": 20261017})
sim = Simulation(bench.write_inp(os.path.join(tempfile.mkdtemp(), "p.inp"), opts), R, 0)
for i in range(8):
    sim.engine.run_async(100, 100000, 0, 1000000)
    sim.engine.exchange_collect(to_host=False)
    sim.exchange_apply(i + 1, None)
sim.engine.synchronize()
w0 = time.perf_counter()
for i in range(3):
    sim.engine.run_async(100, 100000, 0, 1000000)
sim.engine.synchronize()
dt = time.perf_counter() - w0
sim.engine.assert_ok()
print(f"blocks/SM cap {os.environ.get('LDO_MAX_BLOCKS_PER_SM', '-'):>3s}  replicas {R}  {300 * R / dt / 1e6:8.3f} M moves/s  {dt / 3 * 1e3:8.2f} ms/launch", flush=True)
''' % ROOT
R = sys.argv[1] if len(sys.argv) > 1 else "16384"
for cap in [1, 2, 3, 4, 6, 8, 10, 12, 14]:
    env = dict(os.environ, LDO_MAX_BLOCKS_PER_SM=str(cap))
    subprocess.run([sys.executable, "-c", CHILD, R], env=env, check=False)
