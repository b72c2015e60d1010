"""Generates the committed fixtures under tests/golden/ from the reference tree and the oracle build.

Run in the build container only (needs /root/reference and oracle/_ref):
    python tests/golden/make_fixtures.py

inputs/            the reference's example systems / movesets / order-parameter / bias / windows files
                   (data, re-serialised), so that nothing at test time reads /root/reference
inputs/moveset_linker.json, inputs/moveset_linker_heavy.json  hand-written (not from the reference): movesets with the
                   three transform / linker movetypes, option names as read by simulation.cpp:444-513
inputs/moveset_ctcb.json  hand-written (not from the reference): the standard moveset with the two CTRG scaffold moves
                   replaced by CTCBScaffoldRegrowth / CTCBJumpScaffoldRegrowth, same schema as examples/moveset_standard.json
replay_*.npz       value-level RNG tapes recorded from the UNMODIFIED reference (oracle/_ref) with the
                   lattice state, counters and energy after every chunk of MC steps
energies.json      energies / counters / enthalpy-entropy split / pair-energy tables of reference systems
enum_*.json        exact-enumeration weights produced by the reference CLI (simulation_type=enumerate)
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_ref as o  # noqa: E402

REF_EX = "/root/reference/examples"
INPUTS = os.path.join(HERE, "inputs")


def copy_inputs():
    os.makedirs(INPUTS, exist_ok=True)
    for name in ["snodin_unbound.json", "snodin_assembled.json", "four_unbound.json", "moveset_standard.json",
                 "moveset_four.json", "ops_standard.json", "biases_mwus-numfulldomains.json"]:
        with open(os.path.join(REF_EX, name)) as f:
            data = json.load(f)
        with open(os.path.join(INPUTS, name), "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
            f.write("\n")
    with open(os.path.join(REF_EX, "snodin-numfulldomains.windows")) as f:
        text = f.read()
    with open(os.path.join(INPUTS, "snodin-numfulldomains.windows"), "w") as f:
        f.write(text)


def opts(system, moveset, **kw):
    d = dict(o.SNODIN_OPTIONS,
             origami_input_filename=os.path.join(INPUTS, system),
             order_parameter_file=os.path.join(INPUTS, "ops_standard.json"),
             movetype_file=os.path.join(INPUTS, moveset))
    d.update(kw)
    return d


def record_replay(name, options, seed, chunk, n_chunks):
    r = o.RefSystem(options)
    r.seed(seed)
    tapes, lens, states, energies, counters, nchains = [], [], [], [], [], []
    for _ in range(n_chunks):
        r.tape(clear=True)
        r.simulate(chunk)
        tp = r.tape(clear=True)
        tapes.append(tp)
        lens.append(len(tp))
        st = r.state()
        states.append(st)
        energies.append(r.energy())
        counters.append(list(r.counters().values()))
        nchains.append(len(st["chain_index"]))
    maxc = max(nchains)
    maxd = max(len(s["pos"]) for s in states)

    def pad(key, shape, fill=-99):
        out = np.full((n_chunks,) + shape, fill, dtype=np.int32)
        for i, s in enumerate(states):
            a = s[key]
            out[i][tuple(slice(0, n) for n in a.shape)] = a
        return out

    att, acc = r.move_stats()
    np.savez_compressed(
        os.path.join(HERE, f"replay_{name}.npz"),
        tape=np.concatenate(tapes), tape_lens=np.array(lens, dtype=np.int64), chunk=chunk,
        n_chains=np.array(nchains, dtype=np.int32),
        n_domains=np.array([len(s["pos"]) for s in states], dtype=np.int32),
        chain_index=pad("chain_index", (maxc,)), chain_ident=pad("chain_ident", (maxc,)),
        chain_len=pad("chain_len", (maxc,)), pos=pad("pos", (maxd, 3)), ore=pad("ore", (maxd, 3)),
        state=pad("state", (maxd,)), bound=pad("bound", (maxd, 2)),
        energy=np.array(energies), counters=np.array(counters, dtype=np.int32),
        attempts=att, accepts=acc,
        options=json.dumps({k: (os.path.basename(v) if isinstance(v, str) and os.sep in v else v)
                            for k, v in options.items()}))
    print(name, "draws", sum(lens), "moves", att, acc)


def energies_fixture():
    out = {}
    for system, temp in [("snodin_assembled.json", 330), ("snodin_assembled.json", 345), ("snodin_unbound.json", 330),
                         ("four_unbound.json", 330)]:
        r = o.RefSystem(opts(system, "moveset_standard.json", temp=temp), with_sim=False)
        st = r.state()
        idents = set()
        pairs = {}
        with open(os.path.join(INPUTS, system)) as f:
            for chain in json.load(f)["origami"]["identities"]:
                idents.update(chain)
        for a in sorted(idents):
            for b in sorted(idents):
                pe = r.pair_energies(a, b)
                if pe is not None:
                    pairs[f"{a},{b}"] = [float(x) for x in pe[:3]]
        out[f"{system}@{temp}"] = {
            "energy": r.energy(), "counters": r.counters(), "split": {k: float(v) for k, v in r.energy_split().items()},
            "init": [float(x) for x in r.init_energies()], "pair_energies": pairs,
            "n_domains": int(len(st["pos"])),
        }
    with open(os.path.join(HERE, "energies.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")


def enumeration_fixture():
    """Exact enumeration of four_unbound (examples/enum.inp) through the reference CLI."""
    results = {}
    for temp in (330, 340, 345):
        with tempfile.TemporaryDirectory() as tmp:
            options = opts("four_unbound.json", "moveset_four.json", temp=temp, simulation_type="enumerate",
                           min_total_staples=0, max_total_staples=2, max_type_staples=2,
                           enumerate_staples_only=False, output_filebase=os.path.join(tmp, "enum"),
                           ops_to_output="numfulldomains nummisdomains numstackedpairs numstaples")
            inp = os.path.join(tmp, "enum.inp")
            o.write_inp(inp, options)
            res = subprocess.run([o.CLI_PATH, "-i", inp], capture_output=True, text=True, check=True)
            weights = {}
            with open(os.path.join(tmp, "enum.weights")) as f:
                header = f.readline()
                for line in f:
                    parts = line.split()
                    if len(parts) >= 5:
                        weights[" ".join(parts[:4])] = float(parts[4])
            results[str(temp)] = {"header": header.strip(), "weights": weights, "stdout_tail": res.stdout[-600:]}
            print("enum", temp, len(weights), "states")
    with open(os.path.join(HERE, "enum_four_unbound.json"), "w") as f:
        json.dump(results, f, indent=1, sort_keys=True)
        f.write("\n")


REPLAYS = {
    "four_unbound_340K": (("four_unbound.json", "moveset_four.json"), dict(temp=340, max_total_staples=2, max_type_staples=2), 7, 50, 24),
    "snodin_assembled_330K": (("snodin_assembled.json", "moveset_standard.json"), dict(temp=330), 11, 8, 5),
    "snodin_unbound_335K": (("snodin_unbound.json", "moveset_standard.json"), dict(temp=335), 5, 50, 6),
    "snodin_assembled_ctcb_332K": (("snodin_assembled.json", "moveset_ctcb.json"), dict(temp=332), 13, 50, 8),
    "snodin_unbound_ctcb_334K": (("snodin_unbound.json", "moveset_ctcb.json"), dict(temp=334), 17, 60, 6),
    # inputs/moveset_linker*.json are hand-written (the reference ships no moveset with the transform /
    # linker movetypes): CTCBLinkerRegrowth, CTCBClusteredLinkerRegrowth, CTRGLinkerRegrowth
    "snodin_assembled_linker_341K": (("snodin_assembled.json", "moveset_linker_heavy.json"), dict(temp=341), 16, 50, 8),
    "snodin_unbound_linker_336K": (("snodin_unbound.json", "moveset_linker.json"), dict(temp=336), 15, 60, 8),
}

if __name__ == "__main__":
    # python tests/golden/make_fixtures.py [replay names...]: with names, only those replays are re-recorded
    only = sys.argv[1:]
    if not only:
        copy_inputs()
        energies_fixture()
    for name, (files, kw, seed, chunk, n_chunks) in REPLAYS.items():
        if only and name not in only:
            continue
        record_replay(name, opts(*files, **kw), seed=seed, chunk=chunk, n_chunks=n_chunks)
    if not only:
        enumeration_fixture()
