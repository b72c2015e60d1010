"""Operation counts per attempted move, from the host emulation of the device sources (Philox mode): how often a
move calls the expensive primitives, per regime. Guides the optimisation of the run kernel (profiles/README.md)."""
import ctypes
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import conftest  # noqa: E402
from latticednaorigami_b200.binding import Simulation  # noqa: E402

NAMES = ["occupant", "table_put", "table_erase", "bind_domain(compl.)", "check_stacking", "eval_place", "walks_remain_seg",
         "rg_site_lookup", "rg_compute_slot", "rg_fill_feeler_memo", "rg_feeler_general", "step", "slot cache hits", "rg_test_config_avail",
         "recoil slot recomputes", "CTRG moves reaching the weight passes"]
lib = conftest.load_hostsim()
counts = (ctypes.c_longlong * 16).in_dll(lib, "ldo_dbg_counts")
tmp = tempfile.mkdtemp()
for system, temp, moves in [("snodin_assembled.json", 330, 2000), ("snodin_unbound.json", 345, 4000)]:
    sim = Simulation(conftest.write_inp(os.path.join(tmp, f"c{temp}.inp"), conftest.make_options(system, temp=temp, random_seed=3)), 1, 0, lib=lib)
    sim.engine.run(200)
    for i in range(16):
        counts[i] = 0
    a0, _ = sim.engine.move_stats()
    sim.engine.run(moves)
    a1, _ = sim.engine.move_stats()
    print(f"{system} {temp} K, per attempted move ({moves} moves; attempts by movetype {list((a1 - a0)[0])}):")
    print("   " + "  ".join(f"{n} {counts[i] / moves:.1f}" for i, n in enumerate(NAMES)))
