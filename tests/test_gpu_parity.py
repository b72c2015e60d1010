"""Parity tests proper: the CUDA library (through the C-ABI) against tapes recorded from the unmodified
reference, against the live oracle where oracle/_ref travelled, and — at BASELINE sizes — through
size-independent properties (running energy == from-scratch recomputation, full constraint check,
determinism, checkpoint round trip, exact-enumeration statistics)."""
import json
import os

import numpy as np
import pytest

from conftest import (GOLDEN, assert_state_equal, make_options, options_from_fixture, replay_fixture_through, write_inp)
from latticednaorigami_b200.binding import Simulation

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["four_unbound_340K", "snodin_assembled_330K", "snodin_unbound_335K",
                                  "snodin_assembled_ctcb_332K", "snodin_unbound_ctcb_334K",
                                  "snodin_assembled_linker_341K", "snodin_unbound_linker_336K"])
def test_replay_fixture_bit_exact(tmp_path, name):
    fx = np.load(os.path.join(GOLDEN, f"replay_{name}.npz"))
    sim = Simulation(write_inp(str(tmp_path / "r.inp"), options_from_fixture(fx)), 5, 0)
    replay_fixture_through(sim, fx, replicas=[0, 2, 4])


@pytest.mark.parametrize("key", ["snodin_assembled.json@330", "snodin_assembled.json@345", "snodin_unbound.json@330", "four_unbound.json@330"])
def test_energy_of_reference_configurations(tmp_path, key):
    golden = json.load(open(os.path.join(GOLDEN, "energies.json")))[key]
    system, temp = key.split("@")
    sim = Simulation(write_inp(str(tmp_path / "e.inp"), make_options(system, temp=float(temp))), 3, 0)
    e = sim.engine.energies()
    for r in range(3):
        assert abs(e[r, 0] - golden["energy"]) <= 1e-12 * max(1.0, abs(golden["energy"]))
        assert abs(e[r, 1] - golden["split"]["enthalpy"]) <= 1e-12 * max(1.0, abs(golden["split"]["enthalpy"]))
        assert abs(e[r, 2] - golden["split"]["entropy"]) <= 1e-12 * max(1.0, abs(golden["split"]["entropy"]))
        assert abs(e[r, 3] - golden["split"]["stacking"]) <= 1e-9 * max(1.0, abs(golden["split"]["stacking"]))
    keys = ["staples", "domains", "bound_pairs", "fully_bound_pairs", "self_bound_pairs", "misbound_pairs",
            "stacked_pairs", "unassigned", "current_c_i"]
    assert [int(x) for x in sim.engine.counters()[0]] == [golden["counters"][k] for k in keys]


def test_live_replay_against_oracle(oracle, tmp_path):
    cases = [("snodin_assembled.json", 330, 201, 120, {}), ("snodin_unbound.json", 335, 202, 600, {}),
             ("snodin_unbound.json", 345, 203, 600, {}),
             ("four_unbound.json", 330, 204, 2000, {"movetype_file": None, "max_total_staples": 2, "max_type_staples": 2})]
    cases += [("snodin_assembled.json", 336, 205, 800, {"ctcb": True}), ("snodin_unbound.json", 333, 206, 1200, {"ctcb": True})]
    for system, temp, seed, steps, extra in cases:
        opts = make_options(system, "moveset_ctcb.json" if extra.get("ctcb") else "moveset_standard.json", temp=temp)
        if "movetype_file" in extra:
            opts = make_options(system, "moveset_four.json", temp=temp, max_total_staples=2, max_type_staples=2)
        r = oracle.RefSystem(opts)
        r.seed(seed)
        sim = Simulation(write_inp(str(tmp_path / f"{seed}.inp"), opts), 2, 0)
        for _ in range(4):
            r.tape(clear=True)
            r.simulate(steps // 4)
            tape = r.tape(clear=True)
            for rep in (0, 1):
                sim.engine.attach_tape(rep, tape)
            sim.engine.run(steps // 4)
            sim.engine.assert_ok()
            for rep in (0, 1):
                assert sim.engine.tape_position(rep) == len(tape)
                assert_state_equal(sim.engine.state(rep), r.state(), f"{system} seed {seed}")
            e = r.energy()
            assert abs(sim.engine.energies()[0, 0] - e) <= 1e-12 * max(1.0, abs(e))
        att, acc = sim.engine.move_stats()
        ra, rb = r.move_stats()
        assert list(att[0]) == list(ra) and list(acc[0]) == list(rb)


def test_centering_and_constraint_check_replay(oracle, tmp_path):
    opts = make_options("snodin_assembled.json", temp=330, centering_freq=7, constraint_check_freq=5)
    r = oracle.RefSystem(opts)
    r.seed(9)
    r.simulate(60)
    sim = Simulation(write_inp(str(tmp_path / "c.inp"), opts), 1, 0)
    sim.engine.attach_tape(0, r.tape())
    sim.engine.run(60, 7, 0, 5)
    sim.engine.assert_ok()
    assert_state_equal(sim.engine.state(0), r.state())
    assert abs(sim.engine.energies()[0, 0] - r.energy()) <= 1e-12 * abs(r.energy())


def test_philox_ensemble_invariants_at_scale(tmp_path):
    """4096 snodin replicas (BASELINE size): every replica keeps a consistent state."""
    R = 4096
    sim = Simulation(write_inp(str(tmp_path / "p.inp"), make_options("snodin_unbound.json", temp=336, random_seed=1234)), R, 0)
    eng = sim.engine
    eng.run(300, 100, 0, 150)
    eng.assert_ok()
    running = eng.energies()[:, 0]
    recomputed, stacked = eng.recompute_energies()
    assert np.all(np.abs(running - recomputed) <= 1e-9 * np.maximum(1.0, np.abs(recomputed)))
    c = eng.counters()
    assert np.array_equal(stacked, c[:, 6])
    assert np.all(c[:, 7] == 0) and np.all(c[:, 0] >= 0) and np.all(c[:, 0] <= 24)
    assert np.all(c[:, 1] == 24 + 2 * c[:, 0])
    eng.check_all_constraints()
    eng.assert_ok()
    att, acc = eng.move_stats()
    assert np.all(att.sum(axis=1) == 300) and np.all(acc <= att)
    # replicas are independent streams: they must not all have done the same thing
    assert len(np.unique(running)) > R // 4


def test_determinism_and_checkpoint_roundtrip(tmp_path):
    opts = make_options("snodin_assembled.json", temp=332, random_seed=77)
    inp = write_inp(str(tmp_path / "d.inp"), opts)
    a = Simulation(inp, 64, 0)
    b = Simulation(inp, 64, 0)
    a.engine.run(200)
    b.engine.run(80)
    blob = b.engine.checkpoint_save()
    c = Simulation(inp, 64, 0)
    c.engine.checkpoint_load(blob)
    c.engine.run(120)
    b.engine.run(120)
    for eng in (b.engine, c.engine):
        eng.assert_ok()
        assert np.array_equal(eng.energies(), a.engine.energies())
        assert np.array_equal(eng.counters(), a.engine.counters())
    for r in (0, 17, 63):
        assert_state_equal(c.engine.state(r), a.engine.state(r))


def _four_unbound_ensemble(tmp_path, R, burn, sweeps, stride, seed):
    opts = make_options("four_unbound.json", "moveset_four.json", temp=345, max_total_staples=2, max_type_staples=2, random_seed=seed)
    sim = Simulation(write_inp(str(tmp_path / f"s{seed}.inp"), opts), R, 0)
    eng = sim.engine
    idx = [sim.op_tags.index(t) for t in ("numfulldomains", "nummisdomains", "numstackedpairs", "numstaples")]
    eng.run(burn, 1000, 0, 0)
    eng.assert_ok()
    per_rep = {}
    for _ in range(sweeps):
        eng.run(stride, 1000, 0, 0)
        ops = eng.order_params()[:, idx]
        keys, inv = np.unique(ops, axis=0, return_inverse=True)
        for ki, row in enumerate(keys):
            key = "(%d %d %d %d)" % tuple(row)
            per_rep.setdefault(key, np.zeros(R))[np.nonzero(inv.ravel() == ki)[0]] += 1
    eng.assert_ok()
    return {k: v / sweeps for k, v in per_rep.items()}


def test_ensemble_matches_reference_mc_protocol(tmp_path):
    """Production (Philox, lane-parallel) path against the UNMODIFIED reference run with the same short
    protocol on 512 seeds (tests/golden/refmc_four_unbound_345K.json): four_unbound, 345 K, 20000 burn-in
    moves from the unbound start, 40 samples 500 moves apart. Staple-number equilibration is slower than
    this protocol, so both ensembles are compared in the same transient; tolerance 5 combined standard
    errors (replica-to-replica scatter)."""
    ref = json.load(open(os.path.join(GOLDEN, "refmc_four_unbound_345K.json")))
    R = 4096
    freq = _four_unbound_ensemble(tmp_path, R, ref["burn"], ref["samples"], ref["stride"], seed=4242)
    checked = 0
    for key, rv in ref["freq"].items():
        if rv["p"] < 3e-3:
            continue
        per_rep = freq.get(key, np.zeros(R))
        p, sem = per_rep.mean(), per_rep.std(ddof=1) / np.sqrt(R)
        tol = 5 * np.hypot(sem, rv["sem"]) + 1e-4
        assert abs(p - rv["p"]) < tol, (key, p, rv["p"], tol)
        checked += 1
    assert checked >= 5


def test_exact_enumeration_conditional_ratios(tmp_path):
    """Within a fixed staple number the states equilibrate fast, so their RATIOS must already agree with the
    reference's exact enumeration (tests/golden/enum_four_unbound.json, 345 K) after the short protocol."""
    w = json.load(open(os.path.join(GOLDEN, "enum_four_unbound.json")))["345"]["weights"]
    R = 4096
    freq = _four_unbound_ensemble(tmp_path, R, 20000, 40, 500, seed=777)
    for a, b in [("(2 0 1 1)", "(2 0 0 1)"), ("(0 1 0 0)", "(0 0 0 0)")]:
        pa, pb = freq[a], freq[b]
        ratio = pa.mean() / pb.mean()
        rel = np.hypot(pa.std(ddof=1) / np.sqrt(R) / pa.mean(), pb.std(ddof=1) / np.sqrt(R) / pb.mean())
        want = w[a] / w[b]
        assert abs(ratio - want) < 5 * rel * want + 0.03 * want, (a, b, ratio, want, rel)
