"""The command-line driver as a user runs it (apps/main.cpp:19-117 interface): INTEGRATION.md advertises
`latticeDNAOrigami_b200 -i ptmc.inp --replicas 96`."""
import os
import subprocess

import numpy as np
import pytest

from conftest import INPUTS, ROOT, make_options, write_inp

CLI = os.path.join(ROOT, "latticednaorigami_b200", "latticeDNAOrigami_b200")


def ptmc_options(tmp_path, swaps):
    # examples/ptmc.inp with a bounded number of swaps and outputs every 100 steps
    return make_options("snodin_unbound.json", simulation_type="ut_parallel_tempering", swaps=swaps, max_pt_dur=600,
                        exchange_interval=100, num_reps=3, temps=[330.0, 332.0, 334.0], chem_pot_mults=[1, 1, 1],
                        bias_mults=[1, 1, 1], stacking_mults=[1, 1, 1], centering_freq=100000, constraint_check_freq=1000000,
                        output_filebase=str(tmp_path / "ptmc"), configs_output_freq=100, counts_output_freq=100,
                        ops_to_output="numfulldomains nummisdomains numstackedpairs numstaples", order_params_output_freq=100,
                        energies_output_freq=100, random_seed=7)


def test_cli_usage_without_gpu(tmp_path):
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode == 1 and "Input parameter file must be provided" in r.stdout
    r = subprocess.run([CLI, "-i", str(tmp_path / "missing.inp")], capture_output=True, text=True)
    assert r.returncode == 1 and "An exception occurred during the run" in r.stdout


@pytest.mark.gpu
def test_cli_runs_replica_exchange_with_96_replicas(tmp_path):
    swaps = 6
    inp = write_inp(str(tmp_path / "ptmc.inp"), ptmc_options(tmp_path, swaps))
    r = subprocess.run([CLI, "-i", inp, "--replicas", "96"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    # .swp: header of the exchanged control values, then one permutation row per written exchange
    rows = open(tmp_path / "ptmc.swp").read().splitlines()
    assert rows[0].split() == ["330/1/", "332/1/", "334/1/"]
    perms = [[int(x) for x in row.split()] for row in rows[1:]]
    assert len(perms) == swaps + 1 and all(sorted(p) == [0, 1, 2] for p in perms)
    # 96 replicas = 32 ladders x 3 temperatures: one file set per replica, each .trj with swaps frames
    for rep in (0, 1, 50, 95):
        trj = open(tmp_path / f"ptmc-{rep}.trj").read().split("\n\n")
        frames = [f for f in trj if f.strip()]
        assert len(frames) == swaps
        lines = frames[-1].strip().splitlines()
        assert int(lines[0]) == swaps * 100
        idx, ident = (int(x) for x in lines[1].split())
        assert (idx, ident) == (0, 0)
        assert len(lines[2].split()) == 24 * 3 and len(lines[3].split()) == 24 * 3
        ops = np.loadtxt(tmp_path / f"ptmc-{rep}.ops", skiprows=1)
        assert ops.shape == (swaps, 5) and ops[-1, 0] == swaps * 100
        ene = np.loadtxt(tmp_path / f"ptmc-{rep}.ene", skiprows=1)
        assert ene.shape == (swaps, 6)
    assert os.path.exists(tmp_path / "ptmc-95.moves")


@pytest.mark.gpu
def test_cli_constant_temperature_batch(tmp_path):
    opts = make_options("snodin_unbound.json", temp=340, ct_steps=300, output_filebase=str(tmp_path / "ct"), configs_output_freq=100,
                        random_seed=3)
    inp = write_inp(str(tmp_path / "ct.inp"), opts)
    r = subprocess.run([CLI, "-i", inp, "--replicas", "8"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    frames = [f for f in open(tmp_path / "ct-7.trj").read().split("\n\n") if f.strip()]
    assert len(frames) == 3
