"""Upper bound of what gang-scheduling the warps of an SM could give: the run kernel on an ensemble whose
replicas are IDENTICAL (same Philox subsequence, same state, same temperature), so that all resident warps
execute the same instruction stream almost in lock-step and share instruction-cache lines, against the
same ensemble with independent streams. Constant-T snodin, assembled start. Not a benchmark."""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402,F401

import conftest  # noqa: E402
from latticednaorigami_b200.binding import Simulation, _ptr  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for system, temp in [("snodin_assembled.json", 330), ("snodin_unbound.json", 345)]:
    for identical in (False, True):
        opts = conftest.make_options(system, temp=temp, random_seed=7)
        sim = Simulation(conftest.write_inp(os.path.join(tempfile.mkdtemp(), "p.inp"), opts), R, 0)
        eng = sim.engine
        if identical:
            sub = np.zeros(R, dtype=np.uint32)
            eng._check(eng.L.ldo_seed_subsequences(eng.h, 7, _ptr(sub)))
        eng.run(200, 0, 0, 0)
        t0 = time.perf_counter()
        for _ in range(3):
            eng.run_async(100, 0, 0, 0)
        eng.synchronize()
        dt = time.perf_counter() - t0
        eng.assert_ok()
        att, acc = eng.move_stats()
        print(f"{system:24s} T={temp} identical={identical!s:5s} {300 * R / dt / 1e6:8.3f} M moves/s  accepted {acc.sum() / att.sum():.3f}", flush=True)
        del sim
