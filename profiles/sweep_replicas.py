"""Throughput of the run kernel against the number of resident replicas (= warps): tells whether the
instruction-fetch stall is latency-like (throughput scales with warps) or bandwidth-like (flat).
Not a benchmark (see bench.py)."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from latticednaorigami_b200.binding import Simulation  # noqa: E402

L = len(bench.LADDER)
tmp = tempfile.mkdtemp()
for R in [int(x) for x in (sys.argv[1:] or [148, 296, 592, 1184, 2368, 4096, 8192])]:
    R = max(L, R // L * L)
    opts = bench.base_options()
    opts.update({"simulation_type": "ut_parallel_tempering", "num_reps": L, "temps": bench.LADDER, "chem_pot_mults": [1] * L,
                 "bias_mults": [1] * L, "stacking_mults": [1] * L, "exchange_interval": 100, "swaps": 0, "random_seed": 20261017})
    sim = Simulation(bench.write_inp(os.path.join(tmp, f"p{R}.inp"), opts), R, 0)
    for i in range(12):
        sim.engine.run_async(100, 100000, 0, 1000000)
        sim.engine.exchange_collect(to_host=False)
        sim.exchange_apply(i + 1, None)
    sim.engine.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    import time
    w0 = time.perf_counter()
    for i in range(6):
        sim.engine.run_async(100, 100000, 0, 1000000)
    sim.engine.synchronize()
    dt = time.perf_counter() - w0
    sim.engine.assert_ok()
    print(f"replicas {R:6d}  warps/SM {R / 148:6.2f}  {600 * R / dt / 1e6:8.3f} M moves/s  {dt / 6 * 1e3:8.2f} ms/launch", flush=True)
    del sim
