"""Equilibrated 32-temperature snodin ladder for bench.py: the stationary starting ensemble of BOTH bench arms.

The unmodified reference's own replica-exchange driver (UTPTGCMCSimulation, ptmc_simulation.cpp:106-150; oracle
build, one thread per rank over the thread-backed boost::mpi shim) runs examples/ptmc.inp generalised to the bench
ladder - 32 temperatures 330..361 K, exchange_interval 100, moveset_standard, start snodin_unbound - for
SWAPS x 100 = 200 000 moves per replica. The configuration sitting in every temperature slot at the end is written
to bench_data/snodin_ladder32.json (chains in the reference's wire format). bench.py tiles these 32 slot states
over its ensembles (distinct Philox subsequences + a decorrelation warm-up); the reference arm restarts its
processes from the same states (.trj restart files, files.cpp:129-218).

    python bench_data/make_ladder.py          # ~10 min on 8 cores
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

SWAPS, INTERVAL = 2000, 100
LADDER = [330.0 + i for i in range(32)]

if __name__ == "__main__":
    import tempfile

    import oracle_ref
    from bench import base_options
    opts = base_options()
    L = len(LADDER)
    with tempfile.TemporaryDirectory(prefix="ldo_ladder_") as tmp:
        opts.update({"simulation_type": "ut_parallel_tempering", "num_reps": L, "temps": LADDER, "chem_pot_mults": [1] * L,
                     "bias_mults": [1] * L, "stacking_mults": [1] * L, "exchange_interval": INTERVAL, "swaps": SWAPS, "max_pt_dur": 1e9,
                     "restart_from_swap": "false", "configs_output_freq": SWAPS * INTERVAL, "logging_freq": 0,
                     "output_filebase": os.path.join(tmp, "ladder")})
        t0 = time.time()
        res = oracle_ref.pt_run(opts, L, [7000 + 13 * r for r in range(L)], record_tapes=False, workdir=tmp)
        wall = time.time() - t0
    q2r = res["swp"][-1]  # slot -> replica after the last exchange
    slots = []
    for k in range(L):
        st = res["states"][q2r[k]]
        slots.append({"temp": LADDER[k], "chain_index": st["chain_index"].tolist(), "chain_ident": st["chain_ident"].tolist(),
                      "chain_len": st["chain_len"].tolist(), "pos": st["pos"].reshape(-1).tolist(), "ore": st["ore"].reshape(-1).tolist()})
    out = {"generator": "bench_data/make_ladder.py: unmodified reference UTPTGCMCSimulation, oracle build",
           "moves_per_replica": SWAPS * INTERVAL, "exchange_interval": INTERVAL, "ladder": LADDER, "slots": slots,
           "staples_per_slot": [len(s["chain_index"]) - 1 for s in slots], "wall_s": round(wall, 1)}
    with open(os.path.join(HERE, "snodin_ladder32.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
        f.write("\n")
    print("staples per slot:", out["staples_per_slot"], "wall", wall)
