"""GPU call 29 (the last seconds of the round): the serial branch of the FourBody evaluation under replay (a campaign case
that needs the reference's term-by-term sums) and the lane-parallel branch in production mode (running energy against a
recomputation, full constraint check). LDO_CHECK_LIB=hostsim runs the same on the host emulation."""
import os, sys, time
t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tempfile
from pathlib import Path
import numpy as np
import conftest, oracle_ref
from conftest import make_options, write_inp
from latticednaorigami_b200.binding import Simulation
import test_synthetic_systems as ts
lib = conftest.load_hostsim() if os.environ.get("LDO_CHECK_LIB") == "hostsim" else None
tmp = Path(tempfile.mkdtemp())
ts._replay(oracle_ref, make_options("snodin_assembled.json", "moveset_standard.json", temp=348), tmp, lib, seed=3048, steps=300, chunks=6)
print("replay ok", round(time.time() - t0, 2), flush=True)
sim = Simulation(write_inp(str(tmp / "p.inp"), make_options("snodin_assembled.json", temp=336, random_seed=11)), 64, 0, lib=lib)
eng = sim.engine
eng.run(300)
eng.assert_ok()
run_e = eng.energies()[:, 0]
rec, _ = eng.recompute_energies()
assert np.all(np.abs(run_e - rec) <= 1e-9 * np.maximum(1.0, np.abs(rec))), np.abs(run_e - rec).max()
eng.check_all_constraints()
eng.assert_ok()
att, acc = eng.move_stats()
print("production ok", att.sum(axis=0), acc.sum(axis=0), round(float(run_e.mean()), 3), round(time.time() - t0, 2))
