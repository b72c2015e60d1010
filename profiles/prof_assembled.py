"""Profiling driver for the expensive regime: constant-T snodin replicas started from the assembled
configuration at 330 K (every scaffold domain bound, CTRG regrowth through a crowded lattice). Warm-up,
then `--moves` MC moves per replica inside a cudaProfilerStart/Stop window. Not a benchmark."""
import argparse
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import conftest  # noqa: E402
from latticednaorigami_b200.binding import Simulation  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--moves", type=int, default=20)
ap.add_argument("--replicas", type=int, default=4144)
ap.add_argument("--temp", type=float, default=330)
args = ap.parse_args()
opts = conftest.make_options("snodin_assembled.json", temp=args.temp, random_seed=7)
sim = Simulation(conftest.write_inp(os.path.join(tempfile.mkdtemp(), "p.inp"), opts), args.replicas, 0)
sim.engine.run(200, 0, 0, 0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
sim.engine.run(args.moves, 0, 0, 0)
torch.cuda.profiler.stop()
sim.engine.assert_ok()
att, acc = sim.engine.move_stats()
print("moves", att.sum(axis=0), acc.sum(axis=0), "staples", sim.engine.counters()[:, 0].mean())
