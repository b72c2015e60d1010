"""Relaxation of the four_unbound ensemble towards the exact-enumeration distribution: P(numstaples = k) and the
largest state weights against time, to size the burn-in of tests/test_production_parity.py's chi-square test."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import conftest  # noqa: E402
from latticednaorigami_b200.binding import Simulation  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
total = int(sys.argv[2]) if len(sys.argv) > 2 else 300000
stride = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
w = json.load(open(os.path.join(conftest.GOLDEN, "enum_four_unbound.json")))
tmp = tempfile.mkdtemp()
for temp in (330, 340, 345):
    ws = w[str(temp)]["weights"]
    enum_ns = [sum(v for k, v in ws.items() if k.endswith(f" {n})")) for n in range(3)]
    opts = conftest.make_options("four_unbound.json", "moveset_four.json", temp=temp, max_total_staples=2, max_type_staples=2, random_seed=5)
    sim = Simulation(conftest.write_inp(os.path.join(tmp, f"r{temp}.inp"), opts), R, 0)
    t0 = time.perf_counter()
    print(f"T={temp} enum P(staples=0,1,2) = {enum_ns}", flush=True)
    for k in range(total // stride):
        sim.engine.run(stride, 10000, 0, 0)
        c = sim.engine.counters()[:, 0]
        print(f"  moves {(k + 1) * stride:8d}  P(staples) = {[round(float((c == n).mean()), 4) for n in range(3)]}", flush=True)
    dt = time.perf_counter() - t0
    print(f"  {R * total / dt / 1e6:.2f} M moves/s", flush=True)
    sim.close()
