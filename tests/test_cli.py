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


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
def test_cli_replica_exchange_over_two_gpus_matches_one_gpu(tmp_path):
    """--gpus 2: one host thread per GPU, ladders sharded in serpentine order, exchange records all-gathered by NCCL inside
    the C++ host (ldo_sim_exchange_round). Decisions come from a Philox stream shared by the ranks and every replica keeps
    its stream wherever it lives, so the run must reproduce the single-GPU run: same .swp, same trajectories."""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    swaps = 5
    outs = {}
    for gpus in (1, 2):
        d = tmp_path / f"g{gpus}"
        d.mkdir()
        opts = ptmc_options(d, swaps)
        opts.update(num_reps=4, temps=[330.0, 333.0, 336.0, 339.0], chem_pot_mults=[1] * 4, bias_mults=[1] * 4, stacking_mults=[1] * 4)
        inp = write_inp(str(d / "ptmc.inp"), opts)
        r = subprocess.run([CLI, "-i", inp, "--replicas", "64", "--gpus", str(gpus)], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout + r.stderr
        outs[gpus] = d
    assert (outs[1] / "ptmc.swp").read_text() == (outs[2] / "ptmc.swp").read_text()
    # replica k of ladder l: file index l * 4 + k on one GPU; on two GPUs rank (k % 2, reversed in odd pairs) at local
    # index l * 2 + k // 2, file index rank * 32 + local index
    for l in (0, 7, 15):
        for k in range(4):
            rank = k % 2 if (k // 2) % 2 == 0 else 1 - k % 2
            one = (outs[1] / f"ptmc-{l * 4 + k}.trj").read_text()
            two = (outs[2] / f"ptmc-{rank * 32 + l * 2 + k // 2}.trj").read_text()
            assert one == two, (l, k)


@pytest.mark.gpu
def test_cli_enumerates_like_the_reference(tmp_path):
    """examples/enum.inp through the CLI: the reference's stdout summary and .weights (fixture of the reference's run)."""
    import json

    from conftest import GOLDEN
    opts = make_options("four_unbound.json", temp=340, simulation_type="enumerate", min_total_staples=0, max_total_staples=2,
                        max_type_staples=2, enumerate_staples_only=False, output_filebase=str(tmp_path / "enum"),
                        ops_to_output="numfulldomains nummisdomains numstackedpairs numstaples")
    inp = write_inp(str(tmp_path / "enum.inp"), opts)
    r = subprocess.run([CLI, "-i", inp, "--replicas", "2072"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    fx = json.load(open(os.path.join(GOLDEN, "enum_four_unbound.json")))["340"]
    # the tail of the reference's stdout: growthpoint sets of the last staple set, number of configurations, averages
    assert r.stdout.strip().split("\n")[-6:] == fx["stdout_tail"].strip().split("\n")[-6:]
    mine = {}
    for line in open(tmp_path / "enum.weights").read().splitlines()[1:]:
        if line.strip():
            key, value = line.rsplit(" ", 1)
            mine[key] = float(value)
    assert set(mine) == set(fx["weights"])
    for key, value in fx["weights"].items():
        assert mine[key] == pytest.approx(value, rel=2e-5)
