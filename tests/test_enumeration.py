"""Exact enumeration (enumerate.cpp): simulation_type=enumerate through the C++ host and the
device enumerator against the unmodified reference CLI run here on the same inputs (CPU tests: host emulation, live
oracle) and against the committed fixture of the reference's enumeration of examples/enum.inp (GPU test; the reference
needs 22 s per temperature for it). The reference prints six significant digits; everything is compared at that precision:
every state's normalised weight, the number of configurations and the average energy."""
import json
import os
import subprocess

import pytest

from conftest import GOLDEN, INPUTS, make_options, write_inp
from latticednaorigami_b200.binding import Simulation

OPS = "numfulldomains nummisdomains numstackedpairs numstaples"
PRINT_TOL = 2e-5  # half a unit in the sixth significant digit, both sides rounded


def read_weights(path):
    weights = {}
    with open(path) as f:
        header = f.readline().split()
        for line in f:
            parts = line.replace("(", "").replace(")", "").split()
            if len(parts) >= 2:
                weights[tuple(int(x) for x in parts[:-1])] = float(parts[-1])
    return header, weights


def options(tmp, name, **kw):
    base = dict(simulation_type="enumerate", min_total_staples=0, max_total_staples=1, max_type_staples=2,
                enumerate_staples_only=False, ops_to_output=OPS, output_filebase=os.path.join(str(tmp), name))
    base.update(kw)
    system = base.pop("system", "four_unbound.json")
    return write_inp(os.path.join(str(tmp), name + ".inp"), make_options(system, **base))


def reference_enumeration(tmp, **kw):
    from oracle import oracle_ref as o
    res = subprocess.run([o.CLI_PATH, "-i", options(tmp, "ref", **kw)], capture_output=True, text=True, check=True)
    tail = res.stdout.strip().split("\n")
    return read_weights(os.path.join(str(tmp), "ref.weights")), float(tail[-4]), float(tail[-2]), float(tail[-1])


def compare(mine, summary, ref, ref_configs, ref_energy, ref_bias):
    (mh, mw), (rh, rw) = mine, ref
    assert mh == rh
    assert set(mw) == set(rw)
    for key in rw:
        assert mw[key] == pytest.approx(rw[key], rel=PRINT_TOL), key
    assert summary["num_configs"] == pytest.approx(ref_configs, rel=PRINT_TOL)
    assert summary["average_energy"] == pytest.approx(ref_energy, rel=PRINT_TOL, abs=1e-6)
    assert summary["average_bias"] == pytest.approx(ref_bias, rel=PRINT_TOL, abs=1e-6)


VARIANTS = {
    "one_staple": dict(temp=340),
    "disallowed_misbinding": dict(temp=330, misbinding_pot="Disallowed"),
    "mean_field": dict(temp=345, apply_mean_field_cor=True),
    "staples_required_one_per_type": dict(temp=340, min_total_staples=1, max_type_staples=1),
    "three_quarter_turn": dict(temp=335, domain_type="ThreeQuarterTurn"),
    "two_order_parameters": dict(temp=350, ops_to_output="numstaples numfulldomains"),
    "two_staples_disallowed": dict(temp=340, max_total_staples=2, misbinding_pot="Disallowed"),
    "excluded_staple": dict(temp=338, excluded_staples="2"),
    # StapleConformationalEnumerator (enumerate.cpp:666-811): the scaffold keeps the input's configuration
    "staples_only": dict(temp=340, enumerate_staples_only=True, max_total_staples=2),
    "staples_only_disallowed": dict(temp=330, enumerate_staples_only=True, max_total_staples=2, misbinding_pot="Disallowed"),
    "staples_only_on_assembled_snodin": dict(system="snodin_assembled.json", temp=335, enumerate_staples_only=True, max_total_staples=1,
                                             max_type_staples=1),
}


@pytest.mark.parametrize("name", sorted(VARIANTS))
@pytest.mark.parametrize("workers", [1, 5])
def test_enumeration_matches_reference_cli(hostsim_lib, oracle, tmp_path, name, workers):
    """One worker walks the whole recursion tree; five take prefixes of it - same sums."""
    kw = VARIANTS[name]
    ref = reference_enumeration(tmp_path, **kw)
    sim = Simulation(options(tmp_path, "mine", **kw), workers, 0, lib=hostsim_lib)
    sim.run()
    compare(read_weights(os.path.join(str(tmp_path), "mine.weights")), sim.enumeration_summary(), *ref)


def test_enumeration_with_a_bias(hostsim_lib, oracle, tmp_path):
    """calc_and_save_weights adds the move-update biases of the conformation (enumerate.cpp:642-648)."""
    biases = {"origami": {"bias_functions": [
        {"label": "well", "type": "SquareWell", "tag": "sw-full", "ops": ["numfulldomains"], "min_op": 1, "max_op": 2,
         "well_bias": -1.5, "outside_bias": 0.75, "update_per_domain": False}]}}
    bias_path = os.path.join(str(tmp_path), "biases.json")
    with open(bias_path, "w") as f:
        json.dump(biases, f)
    kw = dict(temp=342, bias_functions_file=bias_path)
    ref = reference_enumeration(tmp_path, **kw)
    assert ref[3] != 0
    sim = Simulation(options(tmp_path, "mine", **kw), 3, 0, lib=hostsim_lib)
    sim.run()
    compare(read_weights(os.path.join(str(tmp_path), "mine.weights")), sim.enumeration_summary(), *ref)


def test_prefixes_partition_the_tree(hostsim_lib, tmp_path):
    """Whatever the number of workers (hence the depth at which the tree is cut), every conformation is visited once."""
    leaves = set()
    for workers in (1, 2, 40, 1500):
        sim = Simulation(options(tmp_path, f"w{workers}", temp=340), workers, 0, lib=hostsim_lib)
        sim.run()
        s = sim.enumeration_summary()
        leaves.add((s["leaves"], round(s["num_configs"])))
    assert len(leaves) == 1, leaves


def fixture_case(temp):
    with open(os.path.join(GOLDEN, "enum_four_unbound.json")) as f:
        fx = json.load(f)[str(temp)]
    weights = {tuple(int(x) for x in k.strip("()").split()): v for k, v in fx["weights"].items()}
    tail = fx["stdout_tail"].strip().split("\n")
    return (fx["header"].split(), weights), float(tail[-4]), float(tail[-2]), float(tail[-1])


def test_full_enumeration_matches_fixture_on_host_emulation(hostsim_lib, tmp_path):
    """examples/enum.inp (up to two staples, 40 growthpoint sets, 3.9e9 configurations) at one temperature."""
    sim = Simulation(options(tmp_path, "full", temp=340, max_total_staples=2), 2, 0, lib=hostsim_lib)
    sim.run()
    compare(read_weights(os.path.join(str(tmp_path), "full.weights")), sim.enumeration_summary(), *fixture_case(340))


@pytest.mark.gpu
@pytest.mark.parametrize("temp", [330, 340, 345])
def test_full_enumeration_matches_fixture_on_gpu(tmp_path, temp):
    import time
    sim = Simulation(options(tmp_path, "gpu", temp=temp, max_total_staples=2), 4144, 0)
    t0 = time.time()
    sim.run()
    print(f"enumeration of examples/enum.inp at {temp} K on 4144 workers: {time.time() - t0:.2f} s (reference: 22 s on one core)")
    compare(read_weights(os.path.join(str(tmp_path), "gpu.weights")), sim.enumeration_summary(), *fixture_case(temp))


@pytest.mark.gpu
def test_gpu_enumeration_matches_reference_cli_variants(tmp_path):
    """GPU against the host emulation on the variants of the CPU test (same code, 4144 workers, deeper cut)."""
    from conftest import load_hostsim
    for name in ("mean_field", "two_staples_disallowed", "three_quarter_turn", "staples_only", "staples_only_on_assembled_snodin"):
        kw = VARIANTS[name]
        host = Simulation(options(tmp_path, "h_" + name, **kw), 1, 0, lib=load_hostsim())
        host.run()
        gpu = Simulation(options(tmp_path, "g_" + name, **kw), 2072, 0)
        gpu.run()
        hs, gs = host.enumeration_summary(), gpu.enumeration_summary()
        assert hs["leaves"] == gs["leaves"]
        assert gs["num_configs"] == pytest.approx(hs["num_configs"], rel=1e-12)
        assert gs["average_energy"] == pytest.approx(hs["average_energy"], rel=1e-9)
        (_, hw), (_, gw) = (read_weights(os.path.join(str(tmp_path), f"{p}_{name}.weights")) for p in "hg")
        assert set(hw) == set(gw)
        for key in hw:
            assert gw[key] == pytest.approx(hw[key], rel=PRINT_TOL)
