"""Host-side table builder and parameter-file reader (C++ host, GPU-free entry points) against
  * the reference's own known answers (tests/src/test_nearest_neighbour.cpp:21-24,55-59,83-115,135;
    tests/src/test_ideal_random_walk.cpp:16-45),
  * tables dumped from the unmodified reference (tests/golden/energies.json), 1e-12 relative."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, INPUTS, make_options, write_inp
from latticednaorigami_b200 import binding

REF_SEQS = ["AT", "TA", "TGCA", "CT", "CG", "AATTACAGTCTGACGGCC"]
REF_ENERGIES = [2.5328725104506113, 2.9857702441073424, -5.4850947742870915, 0.9057954673134645,
                -0.18451389148978148, -43.193024598743925]


def test_nn_known_answers():
    for seq, want in zip(REF_SEQS, REF_ENERGIES):
        h, s = binding.host_nn_unitless_thermo(seq, 300, 1.0)
        assert h - s == pytest.approx(want, rel=1e-12)
    h, s = binding.host_nn_unitless_thermo("AT", 300, 0.5)
    assert h - s == pytest.approx(2.7895932252887095, rel=1e-12)


def test_longest_contig_complement_known_answers():
    f = binding.host_longest_contig_complement
    assert f("ATCGAAAAAAAAACTAA", "TTAGAAAAACGATAAAA") == ["ATCG", "CTAA"]
    assert f("TTAGAAAAACGATAAAA", "ATCGAAAAAAAAACTAA") == ["TTAG", "CGAT"]
    assert f("CCTTTTTTTCTTTATA", "TCGCTTCCTACTCCCA") == ["TA", "TA"]
    assert f("TCGCTTCCTACTCCCA", "CCTTTTTTTCTTTATA") == ["TA", "TA"]


def test_walk_predicate_known_answers():
    # num_walks((0,0,0),(0,2,3),5)=10, N=7 -> 665, N=6 -> 0, N=51 -> 5.9e36 (test_ideal_random_walk.cpp:16-45)
    assert not binding.host_no_walks((0, 0, 0), (0, 2, 3), 5)
    assert not binding.host_no_walks((0, 0, 0), (0, 2, 3), 7)
    assert binding.host_no_walks((0, 0, 0), (0, 2, 3), 6)
    assert not binding.host_no_walks((0, 0, 0), (0, 2, 3), 51)
    assert binding.host_no_walks((0, 0, 0), (0, 2, 3), 4)
    assert not binding.host_no_walks((1, 2, 3), (3, 0, 5), 6)  # sign / permutation symmetry
    assert not binding.host_no_walks((3, 0, 5), (1, 2, 3), 6)


@pytest.mark.parametrize("key", ["snodin_assembled.json@330", "snodin_assembled.json@345", "four_unbound.json@330"])
def test_energy_tables_match_reference(tmp_path, key):
    golden = json.load(open(os.path.join(GOLDEN, "energies.json")))[key]
    system, temp = key.split("@")
    inp = write_inp(str(tmp_path / "t.inp"), make_options(system, temp=float(temp)))
    t = binding.host_energy_tables(inp, float(temp))
    n = t["n_ident"]
    assert np.allclose(t["init"], golden["init"], rtol=1e-14, atol=0)
    count = 0
    for pair, (e, h, s) in golden["pair_energies"].items():
        a, b = (int(x) for x in pair.split(","))
        k = (a + n) * (2 * n + 1) + (b + n)
        assert t["present"][k]
        for got, want in ((t["energy"][k], e), (t["enthalpy"][k], h), (t["entropy"][k], s)):
            assert abs(got - want) <= 1e-12 * max(1.0, abs(want)), (pair, got, want)
        count += 1
    assert count == int(t["present"].sum())


def test_inp_reader_defaults_and_lists(tmp_path):
    inp = str(tmp_path / "p.inp")
    with open(inp, "w") as f:
        f.write("# comment\norigami_input_filename=x.json\n temps = 330.0 332.0 334.0 \nchem_pot_mults= 1 1 1\n"
                "num_reps=3\nsimulation_type=ut_parallel_tempering\nops_to_output=a b c\nmax_duration=10 # trailing\n"
                "apply_mean_field_cor=true\n")
    v = lambda k: binding.host_inp_value(inp, k)
    assert v("temps") == "330 332 334" and v("chem_pot_mults") == "1 1 1" and v("num_reps") == "3"
    assert v("domain_type") == "Halfturn"  # reference default (parser.cpp:40), not an accepted value
    assert v("max_total_staples") == "999" and v("max_staple_size") == "2" and v("temp") == "300"
    assert v("restart_traj_postfix") == ".trj" and float(v("max_rel_P_diff")) == 0.1
    assert v("ops_to_output") == "a b c" and v("max_duration") == "10" and v("apply_mean_field_cor") == "true"
    assert v("random_seed") == "-1" and v("simulation_type") == "ut_parallel_tempering"


def test_inp_reader_rejects_unknown_option(tmp_path):
    inp = str(tmp_path / "p.inp")
    open(inp, "w").write("no_such_option=1\n")
    with pytest.raises(binding.LdoError, match="unrecognised option"):
        binding.host_inp_value(inp, "temp")
