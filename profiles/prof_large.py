"""BASELINE.json config 5: synthetic 168-domain ThreeQuarterTurn raster scaffold, 84 two-domain
staple types, Uniform hybridization, 4096 replicas per GPU. The replica state (22.9 KB + scratch) does not fit
shared memory, so the same move code runs in place on HBM/L2 (`k_exec_inplace<CapsLarge>`). Reports attempted
moves/s at a few temperatures of the annealing sweep; with --profile one launch is bracketed by
cudaProfilerStart/Stop for `ncu --profile-from-start off`. Not the headline benchmark (see bench.py)."""
import argparse
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import conftest  # noqa: E402
from latticednaorigami_b200.binding import Simulation  # noqa: E402
from synthetic import UNIFORM_OPTIONS, write_raster_system  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--replicas", type=int, default=4096)
ap.add_argument("--moves", type=int, default=50)
ap.add_argument("--temps", type=float, nargs="+", default=[350.0, 335.0, 320.0])
ap.add_argument("--profile", action="store_true")
args = ap.parse_args()
tmp = tempfile.mkdtemp()
system = write_raster_system(os.path.join(tmp, "raster_12x14.json"), 12, 14)
for temp in args.temps:
    opts = conftest.make_options(temp=temp, max_total_staples=168, max_type_staples=2, staple_M=1e-3, random_seed=11, **UNIFORM_OPTIONS)
    opts["origami_input_filename"] = system
    sim = Simulation(conftest.write_inp(os.path.join(tmp, f"big{int(temp)}.inp"), opts), args.replicas, 0)
    eng = sim.engine
    eng.run(100, 50, 0, 0)
    torch.cuda.synchronize()
    if args.profile:
        torch.cuda.profiler.start()
    t0 = time.perf_counter()
    eng.run(args.moves, 0, 0, 0)
    dt = time.perf_counter() - t0
    if args.profile:
        torch.cuda.profiler.stop()
    eng.assert_ok()
    att, acc = eng.move_stats()
    print(f"T={temp:.0f} K  replicas {args.replicas}  {args.moves * args.replicas / dt / 1e3:9.1f} k moves/s  "
          f"{dt * 1e3:8.1f} ms/launch  accepted {acc.sum() / att.sum():.3f}  staples/replica {eng.counters()[:, 0].mean():.1f}  "
          f"state bytes {eng.state_bytes()}", flush=True)
    del sim
    if args.profile:
        break
