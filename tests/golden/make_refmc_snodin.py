"""Reference-MC ensemble statistics of snodin under a fixed short protocol, from the UNMODIFIED reference
(oracle/_ref): the fixture the production (Philox, lane-parallel) path of the CUDA engine is pinned against
(tests/test_production_parity.py).

Five state points: snodin_unbound at 335 / 340 / 345 K and snodin_assembled at 330 / 336 K, moveset_standard,
512 independent seeds each; BURN moves from the start configuration, then SAMPLES samples STRIDE moves apart of
(numstaples, numfulldomains, nummisdomains, numstackedpairs, energy). Recorded per state point: ensemble mean and
standard error (seed-to-seed scatter of the per-seed time averages) of every observable, and of the acceptance
rate of every movetype over the whole run - the most sensitive observable for a wrong recoil-growth weight.
The protocol is far shorter than equilibration: both codes are compared in the same transient.

    python tests/golden/make_refmc_snodin.py        # ~10 min on 8 cores, writes refmc_snodin.json
"""
import json
import os
import sys
from multiprocessing import Pool

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

BURN, SAMPLES, STRIDE, SEEDS = 2000, 20, 200, 512
POINTS = [("snodin_unbound.json", 335), ("snodin_unbound.json", 340), ("snodin_unbound.json", 345),
          ("snodin_assembled.json", 330), ("snodin_assembled.json", 336)]
OBS = ["numstaples", "numfulldomains", "nummisdomains", "numstackedpairs", "energy"]


def work(job):
    import numpy as np

    import conftest
    import oracle_ref as o
    system, temp, seed = job
    r = o.RefSystem(conftest.make_options(system, temp=temp))
    r.seed(seed)
    r.simulate(BURN)
    acc = np.zeros(len(OBS))
    for _ in range(SAMPLES):
        r.simulate(STRIDE)
        r.tape(clear=True)
        c = r.counters()
        acc += [c["staples"], c["fully_bound_pairs"], c["misbound_pairs"], c["stacked_pairs"], r.energy()]
    att, ok = r.move_stats()
    return (acc / SAMPLES).tolist(), att.tolist(), ok.tolist()


if __name__ == "__main__":
    import numpy as np
    out = {"burn": BURN, "samples": SAMPLES, "stride": STRIDE, "seeds": SEEDS, "observables": OBS, "points": {}}
    with Pool(os.cpu_count()) as p:
        for pi, (system, temp) in enumerate(POINTS):
            res = p.map(work, [(system, temp, 100000 * (pi + 1) + s) for s in range(SEEDS)], chunksize=4)
            obs = np.array([r[0] for r in res])
            att = np.array([r[1] for r in res], dtype=float)
            ok = np.array([r[2] for r in res], dtype=float)
            rate = ok / np.maximum(att, 1)
            key = f"{system.replace('.json', '')}@{temp}"
            out["points"][key] = {
                "system": system, "temp": temp,
                "mean": dict(zip(OBS, obs.mean(axis=0).tolist())),
                "sem": dict(zip(OBS, (obs.std(axis=0, ddof=1) / np.sqrt(SEEDS)).tolist())),
                "accept_rate": rate.mean(axis=0).tolist(),
                "accept_rate_sem": (rate.std(axis=0, ddof=1) / np.sqrt(SEEDS)).tolist(),
                "attempt_share": (att.sum(axis=0) / att.sum()).tolist(),
            }
            print(key, json.dumps(out["points"][key]["mean"]), out["points"][key]["accept_rate"], flush=True)
    with open(os.path.join(HERE, "refmc_snodin.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
