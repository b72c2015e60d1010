"""Profiling driver: brings the bench workload (snodin, stationary 32-temperature ladder of bench_data/) to a
decorrelated state, then runs `--moves` MC moves per replica inside a cudaProfilerStart/Stop window so that
`ncu --profile-from-start off` captures exactly the run launch. Not a benchmark (see bench.py)."""
import argparse
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from latticednaorigami_b200.binding import Simulation  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--moves", type=int, default=100)
ap.add_argument("--rounds", type=int, default=6)
ap.add_argument("--replicas", type=int, default=16384)
args = ap.parse_args()
wl = bench.PTMC()
L = len(bench.LADDER)
tmp = tempfile.mkdtemp()
opts = wl.options()
sim = Simulation(bench.write_inp(os.path.join(tmp, "p.inp"), opts), args.replicas, 0)
bench.tile_states(sim.engine, wl.data["slots"], args.replicas // L, L, opts["random_seed"], lambda r: r)
for i in range(args.rounds):
    sim.engine.run_async(100, 100000, 0, 1000000)
    sim.engine.exchange_collect(to_host=False)
    sim.exchange_apply(i + 1, None)
sim.engine.synchronize()
torch.cuda.synchronize()
torch.cuda.profiler.start()
sim.engine.run(args.moves, 100000, 0, 1000000)
torch.cuda.profiler.stop()
sim.engine.assert_ok()
att, acc = sim.engine.move_stats()
print("moves", att.sum(axis=0), acc.sum(axis=0), "staples", sim.engine.counters()[:, 0].mean())
