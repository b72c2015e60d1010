"""Per-replica wall time inside one run launch of the bench workload (ldo_get_run_timing): how much of the
launch is the slowest replica, and whether a second wave of blocks exists. Diagnostics, not a benchmark."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from latticednaorigami_b200.binding import Simulation  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 20
L = len(bench.LADDER)
tmp = tempfile.mkdtemp()
opts = bench.base_options()
opts.update({"simulation_type": "ut_parallel_tempering", "num_reps": L, "temps": bench.LADDER, "chem_pot_mults": [1] * L,
             "bias_mults": [1] * L, "stacking_mults": [1] * L, "exchange_interval": 100, "swaps": 0, "random_seed": 20261017})
sim = Simulation(bench.write_inp(os.path.join(tmp, "p.inp"), opts), R, 0)
for i in range(rounds):
    sim.engine.run_async(100, 100000, 0, 1000000)
    sim.engine.exchange_collect(to_host=False)
    sim.exchange_apply(i + 1, None)
sim.engine.synchronize()
sim.engine.run(100, 100000, 0, 1000000)
t = sim.engine.run_timing()
t0, t1, sm = t[:, 0], t[:, 1], t[:, 2]
start = t0.min()
dur = (t1 - t0) / 1e6
span = (t1.max() - start) / 1e6
print(f"replicas {R}: launch span {span:.1f} ms; per-replica ms: mean {dur.mean():.1f} p50 {np.median(dur):.1f} "
      f"p95 {np.percentile(dur, 95):.1f} max {dur.max():.1f}; sum/(span*R) = {dur.sum() / (span * R):.2f}")
late = (t0 - start) / 1e6 > 1.0
print(f"replicas starting > 1 ms after the first (second wave): {late.sum()}, their start ms: "
      f"{np.sort((t0[late] - start) / 1e6)[:5]} .. end {(t1[late].max() - start) / 1e6 if late.any() else 0:.1f}")
done = np.sort((t1 - start) / 1e6)
print("finish time quantiles ms (10..100%):", [round(float(np.percentile(done, q)), 1) for q in range(10, 101, 10)])
ti = sim.engine.control()["temp_idx"]
for k in range(0, L, 4):
    m = ti == k
    print(f"  T={bench.LADDER[k]:6.1f}  replicas {m.sum():4d}  mean {dur[m].mean():7.1f} ms  max {dur[m].max():7.1f}")
st = sim.engine.counters()[:, 0]
for lo, hi in [(0, 0), (1, 4), (5, 10), (11, 24)]:
    m = (st >= lo) & (st <= hi)
    if m.any():
        print(f"  staples {lo:2d}-{hi:2d}: replicas {m.sum():4d}  mean {dur[m].mean():7.1f} ms  max {dur[m].max():7.1f}")
