"""Inputs the device path refuses, with the reason (no silent fallback): a Sum of per-domain-update order parameters, the
movetype the reference names but never constructs, and linker-move options out of range."""
import json
import os

import pytest

from conftest import INPUTS, make_options, write_inp
from latticednaorigami_b200.binding import LdoError, Simulation


def _moveset(tmp_path, entry):
    path = tmp_path / "ms.json"
    path.write_text(json.dumps({"origami": {"movetypes": [
        {"label": "Orientation rotation", "type": "OrientationRotation", "freq": "1/2"}, dict(entry, freq="1/2")]}}))
    return str(path)


def test_sum_of_per_domain_order_parameters_is_refused(hostsim_lib, tmp_path):
    """Per-domain Dist / AdjacentSite parameters are supported (tests/test_per_domain_biases.py); a Sum built on them is
    not: the reference's result then depends on the order of its candidate checks."""
    ops = json.load(open(os.path.join(INPUTS, "ops_dist.json")))
    for k in (2, 3):
        ops["origami"]["order_params"][k]["update_per_domain"] = True
    path = tmp_path / "ops.json"
    path.write_text(json.dumps(ops))
    opts = make_options("snodin_unbound.json", domain_update_biases_present=True)
    opts["order_parameter_file"] = str(path)
    with pytest.raises(LdoError, match="Sum of per-domain"):
        Simulation(write_inp(str(tmp_path / "a.inp"), opts), 1, 0, lib=hostsim_lib)


def test_unconstructed_reference_movetype_is_refused(hostsim_lib, tmp_path):
    opts = make_options("snodin_unbound.json")
    opts["movetype_file"] = _moveset(tmp_path, {"label": "x", "type": "CTRGClusteredLinkerRegrowth"})
    with pytest.raises(LdoError, match="CTRGClusteredLinkerRegrowth"):
        Simulation(write_inp(str(tmp_path / "b.inp"), opts), 1, 0, lib=hostsim_lib)


def test_linker_options_out_of_range(hostsim_lib, tmp_path):
    opts = make_options("snodin_unbound.json")
    opts["movetype_file"] = _moveset(tmp_path, {
        "label": "x", "type": "CTCBLinkerRegrowth", "max_disp": 1, "max_turns": 1, "max_regrowth": 4,
        "max_linker_length": 3, "num_transforms": 40})
    with pytest.raises(LdoError, match="linker regrowth options"):
        Simulation(write_inp(str(tmp_path / "c.inp"), opts), 1, 0, lib=hostsim_lib)
