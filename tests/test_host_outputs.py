"""Output files of the C++ host (`ldo_sim_run`) against the files the unmodified reference CLI writes for the
same run, byte for byte. The run is chosen so that its trajectory does not depend on the random stream: the
assembled snodin system at 300 K with a moveset of staple exchanges only and `max_total_staples` equal to the
12 staples present — insertions are refused (met_movetypes.cpp:311-318) and a deletion would cost ~40 kT —
so both programs write the same configurations, counts, energies and order parameters at every output step.
Covers .trj .vsf .vcf .states .ores .counts .staples .staplestates .ene .ops (files.cpp:519-793) and the
general part of .moves (movetypes.cpp:87-96; the movetypes' own breakdowns are compared on evolving runs in
test_restart_and_outputs.py)."""
import json
import os
import subprocess

from conftest import make_options, write_inp
from latticednaorigami_b200.binding import Simulation

EXTS = [".trj", ".vsf", ".vcf", ".states", ".ores", ".counts", ".staples", ".staplestates", ".ene", ".ops"]


def _options(tmp_path, base):
    moveset = tmp_path / "exchange_only.json"
    moveset.write_text(json.dumps({"origami": {"movetypes": [
        {"label": "Met staple exchange", "type": "MetStapleExchange", "freq": "1", "adaptive_exchange": False}]}}))
    opts = make_options("snodin_assembled.json", temp=300, max_total_staples=12, ct_steps=40, random_seed=3,
                        configs_output_freq=10, vtf_output_freq=10, counts_output_freq=10, order_params_output_freq=10,
                        energies_output_freq=10, times_output_freq=0, logging_freq=0,
                        ops_to_output="numstaples numfulldomains nummisdomains numstackedpairs",
                        output_filebase=str(tmp_path / base))
    opts["movetype_file"] = str(moveset)
    return opts


def test_output_files_match_reference_cli(hostsim_lib, oracle, tmp_path):
    ref_inp = write_inp(str(tmp_path / "ref.inp"), _options(tmp_path, "ref"))
    subprocess.run([oracle.CLI_PATH, "-i", ref_inp], check=True, capture_output=True)
    sim = Simulation(write_inp(str(tmp_path / "our.inp"), _options(tmp_path, "our")), 1, 0, lib=hostsim_lib)
    sim.run()
    for ext in EXTS:
        ref, our = (tmp_path / ("ref" + ext)).read_text(), (tmp_path / ("our" + ext)).read_text()
        assert ref, ext
        assert our == ref, f"{ext} differs from the reference's file"
    # .trj holds 4 frames of 13 chains: the comparison is not vacuous
    assert (tmp_path / "our.trj").read_text().count("\n\n") == 4
    ref_moves = (tmp_path / "ref.moves").read_text().splitlines()[:4]
    our_moves = (tmp_path / "our.moves").read_text().splitlines()[:4]
    assert our_moves == ref_moves


def test_annealing_output_files_match_reference_cli(hostsim_lib, oracle, tmp_path):
    """The annealing driver (annealing_simulation.cpp:38-49) on the same frozen trajectory: three temperatures,
    energies and order parameters re-evaluated at every temperature step."""
    def options(base):
        opts = _options(tmp_path, base)
        opts.update(simulation_type="annealing", max_temp=302, min_temp=300, temp_interval=1, steps_per_temp=20)
        return opts
    ref_inp = write_inp(str(tmp_path / "ref_a.inp"), options("refa"))
    subprocess.run([oracle.CLI_PATH, "-i", ref_inp], check=True, capture_output=True)
    sim = Simulation(write_inp(str(tmp_path / "our_a.inp"), options("oura")), 1, 0, lib=hostsim_lib)
    sim.run()
    for ext in EXTS:
        ref, our = (tmp_path / ("refa" + ext)).read_text(), (tmp_path / ("oura" + ext)).read_text()
        assert ref, ext
        assert our == ref, f"{ext} differs from the reference's file"
    assert len((tmp_path / "oura.ene").read_text().splitlines()) == 7  # header + 60 steps / 10
