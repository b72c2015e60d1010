"""Replay campaign of the replica-exchange drivers against the reference's own PT drivers (thread-per-rank oracle): the five
variants (ut / t / st / hut / 2d) on snodin unbound and assembled, 8 exchange rounds of 100 moves, fresh seeds; every swap
decision, the .swp sequence and every replica's state (tests/test_exchange_oracle.py::exchange_against_oracle).
python tests/stress_replay_pt.py SECONDS. Round 2, final build: 959 runs in 540 s, 4 reported - two of them a swap decision
that differs from the reference's, two more than two exchange draws consumed differently. All four pass on the build
before the weight passes of recoil growth stopped evaluating the potential: the running energy then carries
the reference's rounding residue bit for bit, and a swap between replicas of EQUAL energy has p = exp(residue) - a draw is
consumed or not depending on the sign of 1e-13, and the next pair of the round reads a shifted tape."""
import sys, os, tempfile, time, pathlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, ROOT)
import conftest, oracle_ref as o
import test_exchange_oracle as t
t_end=time.time()+float(sys.argv[1]); n=0; bad=0; seed0=10000
lib=conftest.load_hostsim()
while time.time()<t_end:
    for case in t.CASES:
        for system in ("snodin_unbound.json","snodin_assembled.json"):
            if time.time()>t_end: break
            seed0+=100; n+=1
            d=pathlib.Path(tempfile.mkdtemp())
            try: t.exchange_against_oracle(o, d, lib, case, system=system, swaps=8, interval=100, seed0=seed0)
            except AssertionError as ex: bad+=1; print(case, system, seed0, "FAIL", str(ex)[:200], flush=True)
            except Exception as ex: bad+=1; print(case, system, seed0, "EXC", type(ex).__name__, str(ex)[:200], flush=True)
print("runs",n,"failures",bad)
