"""Profiling driver: brings the bench workload (4096 snodin replicas, 32-temperature ladder) to a warmed-up
state, then runs `--moves` MC moves per replica inside a cudaProfilerStart/Stop window so that
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
ap.add_argument("--moves", type=int, default=20)
ap.add_argument("--rounds", type=int, default=30)
ap.add_argument("--replicas", type=int, default=4096)
args = ap.parse_args()
L = len(bench.LADDER)
tmp = tempfile.mkdtemp()
opts = bench.base_options()
opts.update({"simulation_type": "ut_parallel_tempering", "num_reps": L, "temps": bench.LADDER, "chem_pot_mults": [1] * L,
             "bias_mults": [1] * L, "stacking_mults": [1] * L, "exchange_interval": 100, "swaps": 0, "random_seed": 20261017})
sim = Simulation(bench.write_inp(os.path.join(tmp, "p.inp"), opts), args.replicas, 0)
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
