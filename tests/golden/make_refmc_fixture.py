"""Reference-MC ensemble statistics under a fixed short protocol (companion of make_fixtures.py).

four_unbound at 345 K, moveset_four, 512 independent seeds of the UNMODIFIED reference (oracle/_ref), each
20000 burn-in moves from the unbound start then 40 samples 500 moves apart of
(numfulldomains, nummisdomains, numstackedpairs, numstaples). Staple-number transitions are slow (SURVEY.md
§8c), so after this protocol the ensemble is NOT yet at the exact-enumeration distribution; the GPU
engine run with the same protocol must reproduce these transient frequencies within sampling error."""
import collections
import json
import os
import sys
from multiprocessing import Pool

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

BURN, SAMPLES, STRIDE, SEEDS = 20000, 40, 500, 512


def work(seed):
    import conftest
    import oracle_ref as o
    opts = conftest.make_options("four_unbound.json", "moveset_four.json", temp=345, max_total_staples=2, max_type_staples=2)
    r = o.RefSystem(opts)
    r.seed(seed)
    r.simulate(BURN)
    out = []
    for _ in range(SAMPLES):
        r.simulate(STRIDE)
        r.tape(clear=True)
        c = r.counters()
        out.append("(%d %d %d %d)" % (c["fully_bound_pairs"], c["misbound_pairs"], c["stacked_pairs"], c["staples"]))
    return out


if __name__ == "__main__":
    import numpy as np
    with Pool(os.cpu_count()) as p:
        res = p.map(work, range(5000, 5000 + SEEDS))
    keys = sorted({k for r in res for k in r})
    freq = {}
    for k in keys:
        per_seed = np.array([sum(1 for x in r if x == k) / SAMPLES for r in res])
        freq[k] = {"p": float(per_seed.mean()), "sem": float(per_seed.std(ddof=1) / np.sqrt(SEEDS))}
    with open(os.path.join(HERE, "refmc_four_unbound_345K.json"), "w") as f:
        json.dump({"burn": BURN, "samples": SAMPLES, "stride": STRIDE, "seeds": SEEDS, "freq": freq}, f, indent=1, sort_keys=True)
        f.write("\n")
    for k, v in sorted(freq.items(), key=lambda kv: -kv[1]["p"])[:8]:
        print(k, v)
