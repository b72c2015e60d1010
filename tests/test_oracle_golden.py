"""Pins the oracle (the unmodified reference built by oracle/Makefile) against the reference's own
known answers and against the committed fixtures. Skipped where oracle/_ref is not built."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, fixture_state, assert_state_equal, make_options, options_from_fixture
from test_host_tables import REF_ENERGIES, REF_SEQS


def test_oracle_nn_known_answers(oracle):
    for seq, want in zip(REF_SEQS, REF_ENERGIES):
        assert oracle.nn_unitless_energy(seq, 300, 1.0) == pytest.approx(want, rel=1e-12)
    assert oracle.nn_unitless_energy("AT", 300, 0.5) == pytest.approx(2.7895932252887095, rel=1e-12)
    assert oracle.nn_longest_contig_complement("ATCGAAAAAAAAACTAA", "TTAGAAAAACGATAAAA") == ["ATCG", "CTAA"]
    assert oracle.nn_longest_contig_complement("CCTTTTTTTCTTTATA", "TCGCTTCCTACTCCCA") == ["TA", "TA"]


def test_oracle_walk_counts_known_answers(oracle):
    assert oracle.num_walks((0, 0, 0), (0, 2, 3), 5) == 10
    assert oracle.num_walks((0, 0, 0), (0, 2, 3), 7) == 665
    assert oracle.num_walks((0, 0, 0), (0, 2, 3), 6) == 0
    assert oracle.num_walks((0, 0, 0), (0, 2, 3), 51) == pytest.approx(5.947398897268465e36, rel=1e-12)


def test_oracle_reference_energy(oracle):
    r = oracle.RefSystem(make_options("snodin_assembled.json"), with_sim=False)
    assert r.energy() == pytest.approx(-492.22455493278574, rel=1e-14)
    c = r.counters()
    assert (c["staples"], c["fully_bound_pairs"], c["misbound_pairs"], c["stacked_pairs"]) == (12, 24, 0, 16)
    golden = json.load(open(os.path.join(GOLDEN, "energies.json")))["snodin_assembled.json@330"]
    assert r.energy() == golden["energy"]


@pytest.mark.parametrize("name", ["four_unbound_340K", "snodin_assembled_330K"])
def test_oracle_replays_committed_tape(oracle, name):
    """The recorded tapes reproduce the recorded states when replayed through the reference itself."""
    fx = np.load(os.path.join(GOLDEN, f"replay_{name}.npz"))
    r = oracle.RefSystem(options_from_fixture(fx))
    r.set_replay(fx["tape"])
    try:
        for i in range(len(fx["tape_lens"])):
            r.simulate(int(fx["chunk"]))
            assert_state_equal(r.state(), fixture_state(fx, i), f"chunk {i}")
            assert r.energy() == float(fx["energy"][i])
    finally:
        r.set_replay(fx["tape"][:0])
