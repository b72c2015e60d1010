"""Replica exchange: the on-device swap rule against a numpy restatement of
PTGCMCSimulation::calc_acceptance_p (ptmc_simulation.cpp:275-313), and the multi-rank path (split
ladders + all-gather of the dependent quantities) under the gloo backend with world_size 2, run on the
host emulation of the device sources (no GPU needed)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, make_options, write_inp
from latticednaorigami_b200.binding import Simulation

TEMPS = [330.0, 332.0, 334.0, 336.0]


def pt_options(**kw):
    return make_options("snodin_unbound.json", simulation_type="ut_parallel_tempering", num_reps=len(TEMPS),
                        temps=TEMPS, chem_pot_mults=[1] * len(TEMPS), bias_mults=[1] * len(TEMPS),
                        stacking_mults=[1] * len(TEMPS), exchange_interval=20, swaps=4, random_seed=99, **kw)


def acceptance_p(t1, t2, d1, d2):
    """calc_acceptance_p with equal multipliers: only the enthalpy / bias / stacking terms remain."""
    DB = 1 / t2 - 1 / t1
    DH = d2[0] * t2 - d1[0] * t1
    Dst = d2[2] * t2 - d1[2] * t1
    DBM = 1 / t2 - 1 / t1
    DBias = d2[1] * t2 - d1[1] * t1
    return min(1.0, np.exp(DB * (DH + DBias) + DBM * Dst))


def test_single_rank_exchange_rule(hostsim_lib, tmp_path):
    n_ladders = 6
    sim = Simulation(write_inp(str(tmp_path / "pt.inp"), pt_options()), n_ladders * len(TEMPS), 0, lib=hostsim_lib)
    L = len(TEMPS)
    q2r_prev = np.tile(np.arange(L, dtype=np.int32), (n_ladders, 1))
    for swap_i in range(1, 7):
        assert sim.exchange_advance() == 0
        dep = sim.engine.exchange_collect().reshape(n_ladders, L, -1)
        sim.exchange_apply(swap_i)
        q2r, att, acc = sim.exchange_state(n_ladders, L)
        ctl = sim.engine.control()["temp_idx"].reshape(n_ladders, L)
        for l in range(n_ladders):
            for i in range(swap_i % 2, L - 1, 2):
                r1, r2 = q2r_prev[l, i], q2r_prev[l, i + 1]
                p = acceptance_p(TEMPS[i], TEMPS[i + 1], dep[l, r1], dep[l, r2])
                swapped = q2r[l, i] == r2 and q2r[l, i + 1] == r1
                if p == 1.0:
                    assert swapped  # p == 1 accepts without a draw (App. A9)
                if p < 1e-12:
                    assert not swapped
            # control variables follow the permutation: replica q2r[i] now runs at slot i
            for i in range(L):
                assert ctl[l, q2r[l, i]] == i
            assert sorted(q2r[l]) == list(range(L))
        q2r_prev = q2r.copy()
        assert att.sum() == sum(len(range(s % 2, L - 1, 2)) for s in range(1, swap_i + 1)) * n_ladders
    sim.engine.assert_ok()
    e = sim.engine.energies()[:, 0]
    re, _ = sim.engine.recompute_energies()
    assert np.allclose(e, re, rtol=1e-10, atol=1e-9)


WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from latticednaorigami_b200.binding import Simulation
import conftest
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
n_ladders, L = 5, 4
S = L // world
sim = Simulation({inp!r}, n_ladders * S, 0, rank=rank, n_ranks=world, lib=conftest.load_hostsim())
for swap_i in range(1, 6):
    assert sim.exchange_advance() == 0
    dep = torch.from_numpy(sim.engine.exchange_collect().copy())
    gathered = [torch.zeros_like(dep) for _ in range(world)]
    dist.all_gather(gathered, dep)
    sim.exchange_apply(swap_i, torch.stack(gathered).numpy())
q2r, att, acc = sim.exchange_state(n_ladders, L, two_d={two_d!r})
np.savez({out!r} + str(rank), q2r=q2r, att=att, acc=acc, energy=sim.engine.energies(), ctl=sim.engine.control()["temp_idx"])
dist.destroy_process_group()
"""


@pytest.mark.parametrize("two_d", [False, True])
def test_two_rank_exchange_matches_single_rank(hostsim_lib, tmp_path, two_d):
    """world_size 2 over gloo: ranks hold half of every ladder, all-gather the dependent quantities and
    must take exactly the decisions of a single rank holding everything (1-D ladder and 2-D
    temperature x stacking-multiplier grid)."""
    if two_d:
        opts = make_options("snodin_unbound.json", simulation_type="2d_parallel_tempering", num_reps=4, temps=[330.0, 336.0],
                            stacking_mults=[1.0, 0.8], exchange_interval=20, swaps=4, random_seed=99)
    else:
        opts = pt_options()
    inp = write_inp(str(tmp_path / "pt.inp"), opts)
    n_ladders, L = 5, len(TEMPS)
    one = Simulation(inp, n_ladders * L, 0, lib=hostsim_lib)
    for swap_i in range(1, 6):
        assert one.exchange_advance() == 0
        one.exchange_apply(swap_i)
    q2r1, att1, acc1 = one.exchange_state(n_ladders, L, two_d=two_d)
    e1 = one.engine.energies().reshape(n_ladders, L, 5)

    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, inp=inp, out=str(tmp_path / "out"), two_d=two_d))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29632" if two_d else "29631", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r))) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    outs = [np.load(str(tmp_path / f"out{r}.npz")) for r in range(2)]
    for o in outs:
        assert np.array_equal(o["q2r"], q2r1) and np.array_equal(o["att"], att1) and np.array_equal(o["acc"], acc1)
    # same trajectories: replica k of ladder l lives on rank (k % 2, reversed in odd groups of 2) at local
    # index l * S + k // 2 (serpentine dealing of the ladder slots)
    S = L // 2
    for l in range(n_ladders):
        for k in range(L):
            rank_of_k = k % 2 if (k // 2) % 2 == 0 else 1 - k % 2
            got = outs[rank_of_k]["energy"][l * S + k // 2]
            assert np.array_equal(got, e1[l, k]), (l, k)


def _two_d_exchange_rule(lib, tmp_path):
    """2d_parallel_tempering (TwoDPTGCMCSimulation, ptmc_simulation.cpp:428-601): temperatures x stacking
    multipliers, four alternating pair sets; decisions checked against a numpy restatement of
    calc_acceptance_p and of the pair schedule (ptmc_simulation.hpp:158-164)."""
    temps, smults = [330.0, 334.0, 338.0], [1.0, 0.8]
    v1, v2 = len(temps), len(smults)
    L = v1 * v2
    n_ladders = 5
    opts = make_options("snodin_unbound.json", simulation_type="2d_parallel_tempering", num_reps=L, temps=temps,
                        stacking_mults=smults, exchange_interval=20, swaps=4, random_seed=77)
    sim = Simulation(write_inp(str(tmp_path / "pt2d.inp"), opts), n_ladders * L, 0, lib=lib)
    ctl = sim.engine.control()
    assert np.array_equal(ctl["temp_idx"].reshape(n_ladders, L)[0], np.repeat(np.arange(v1), v2))
    assert np.allclose(ctl["stacking_mult"].reshape(n_ladders, L)[0], np.tile(smults, v1))
    slot_t = np.repeat(temps, v2)
    slot_sm = np.tile(smults, v1)
    q2r_prev = np.tile(np.arange(L, dtype=np.int32), (n_ladders, 1))
    expected_attempts = np.zeros((2, L), dtype=np.int64)
    starts = {0: (0, 0, 2, 1, v1 - 1, v2, v2), 1: (0, 0, 1, 2, v1, v2 - 1, 1),
              2: (1, 0, 2, 1, v1 - 1, v2, v2), 3: (0, 1, 1, 2, v1, v2 - 1, 1)}
    for swap_i in range(1, 10):
        assert sim.exchange_advance() == 0
        dep = sim.engine.exchange_collect().reshape(n_ladders, L, -1)
        sim.exchange_apply(swap_i)
        q2r, att, acc = sim.exchange_state(n_ladders, L, two_d=True)
        i0, j0, di, dj, i1, j1, incr = starts[swap_i % 4]
        pairs = [(i * v2 + j, i * v2 + j + incr) for i in range(i0, i1, di) for j in range(j0, j1, dj)]
        for a, _ in pairs:
            expected_attempts[swap_i % 2, a] += 1
        for l in range(n_ladders):
            touched = set()
            for a, b in pairs:
                touched.update((a, b))
                r1, r2 = q2r_prev[l, a], q2r_prev[l, b]
                t1, t2 = slot_t[a], slot_t[b]
                d1, d2 = dep[l, r1], dep[l, r2]
                DB = 1 / t2 - 1 / t1
                DH = d2[0] * t2 - d1[0] * t1
                DBias = d2[1] * t2 - d1[1] * t1
                Dst = d2[2] * t2 - d1[2] * t1
                DBM = slot_sm[b] / t2 - slot_sm[a] / t1
                # equal staple_u multipliers: the chemical-potential term vanishes
                p = min(1.0, np.exp(DB * (DH + DBias) + DBM * Dst))
                swapped = q2r[l, a] == r2 and q2r[l, b] == r1
                if p == 1.0:
                    assert swapped
                if p < 1e-12:
                    assert not swapped
                assert swapped or (q2r[l, a] == r1 and q2r[l, b] == r2)
            for k in range(L):
                if k not in touched:
                    assert q2r[l, k] == q2r_prev[l, k]
            assert sorted(q2r[l]) == list(range(L))
            # control variables follow the permutation
            ctl = sim.engine.control()
            for k in range(L):
                r = l * L + q2r[l, k]
                assert ctl["temp_idx"][r] == k // v2 and ctl["stacking_mult"][r] == slot_sm[k]
            assert np.array_equal(att[l], expected_attempts)
        q2r_prev = q2r.copy()
    assert acc.sum() > 0
    sim.engine.assert_ok()
    e = sim.engine.energies()[:, 0]
    re, _ = sim.engine.recompute_energies()
    assert np.allclose(e, re, rtol=1e-10, atol=1e-9)


def test_two_d_exchange_rule(hostsim_lib, tmp_path):
    _two_d_exchange_rule(hostsim_lib, tmp_path)


@pytest.mark.gpu
def test_two_d_exchange_rule_gpu(tmp_path):
    _two_d_exchange_rule(None, tmp_path)


def test_swap_probability_matches_reference_code(hostsim_lib, oracle, tmp_path):
    """The swap probability the exchange kernels evaluate (`exchange_acceptance_p`, through its host export)
    against PTGCMCSimulation::calc_acceptance_p of the unmodified reference, on random control variables and
    dependent quantities including unequal chemical-potential and stacking multipliers; 1e-12 relative. Also
    pins the reduced staple chemical potentials the engine holds."""
    opts = pt_options(output_filebase=str(tmp_path / "ref"))
    ref = oracle.RefSystem(opts, with_sim=False)
    sim = Simulation(write_inp(str(tmp_path / "pt.inp"), pt_options()), len(TEMPS), 0, lib=hostsim_lib)
    eng = sim.engine
    red = np.zeros(64)
    nst = eng.L.ldo_get_reduced_staple_u(eng.h, red.ctypes.data)
    red = red[:nst]
    us = ref.staple_us(341.0, 1.3)
    assert len(us) == nst == 12
    assert np.allclose(us, red * 341.0 * 1.3, rtol=1e-13, atol=0)
    rng = np.random.default_rng(5)
    seen_partial = 0
    for _ in range(200):
        t1, t2 = rng.uniform(320, 360, 2)
        um1, um2 = rng.uniform(0.8, 1.2, 2)
        sm1, sm2 = rng.uniform(0.5, 1.0, 2)
        scale = rng.choice([0.01, 0.3, 3.0])
        # exchange record layout: enthalpy, bias, stacking, the replica's staple_u multiplier, staple counts
        d1 = np.concatenate([rng.normal(-300, 60, 1) * scale, rng.normal(0, 3, 1) * scale, rng.normal(-20, 5, 1) * scale, [um1],
                             rng.integers(0, 3, nst).astype(float)])
        d2 = np.concatenate([rng.normal(-300, 60, 1) * scale, rng.normal(0, 3, 1) * scale, rng.normal(-20, 5, 1) * scale, [um2],
                             rng.integers(0, 3, nst).astype(float)])
        ours = eng.L.ldo_exchange_acceptance_p(nst, red.ctypes.data, t1, t2, um1, um2, sm1, sm2, d1.ctypes.data, d2.ctypes.data)
        pad = lambda a: np.concatenate([a, [0.0]])  # the reference's loop reads n_types entries (App. A1)
        want = oracle.pt_acceptance_p(ref, (t1, t2), (um1, um2), (1.0, 1.0), (sm1, sm2), d1[:3], d2[:3],
                                      pad(ref.staple_us(t1, um1)), pad(ref.staple_us(t2, um2)), pad(d1[4:]), pad(d2[4:]))
        assert abs(ours - want) <= 1e-12 * max(want, 1e-300), (ours, want)
        seen_partial += 0.0 < want < 1.0
    assert seen_partial > 20


def test_two_d_driver_writes_swap_file(hostsim_lib, tmp_path):
    """ldo_sim_run for 2d_parallel_tempering: the .swp header lists temperature / staple_u_mult / stacking_mult of
    every slot (m_exchange_q_is, ptmc_simulation.cpp:92-104,446-448) and every entry is a permutation."""
    temps, smults = [330.0, 336.0], [1.0, 0.8]
    opts = make_options("snodin_unbound.json", simulation_type="2d_parallel_tempering", num_reps=4, temps=temps,
                        stacking_mults=smults, exchange_interval=10, swaps=4, random_seed=5, configs_output_freq=10,
                        max_pt_dur=1e9, output_filebase=str(tmp_path / "pt2d"))
    sim = Simulation(write_inp(str(tmp_path / "pt2d.inp"), opts), 4, 0, lib=hostsim_lib)
    sim.run()
    lines = (tmp_path / "pt2d.swp").read_text().splitlines()
    assert lines[0].split() == ["330/1/1/", "330/1/0.8/", "336/1/1/", "336/1/0.8/"]
    assert len(lines) >= 5
    for row in lines[1:]:
        assert sorted(int(x) for x in row.split()) == [0, 1, 2, 3]
    # per-replica configuration files carry the "-<rank>" postfix of the reference
    assert (tmp_path / "pt2d-0.trj").exists() and (tmp_path / "pt2d-3.trj").exists()


def _swap_acceptance_frequency(lib, tmp_path, n_ladders, rounds):
    """Decisions against probabilities over many ladders: every tested pair is a Bernoulli trial whose probability
    is the engine's own exchange_acceptance_p (pinned to the reference's calc_acceptance_p above) of the collected
    exchange records. The number of accepted swaps must match the sum of the probabilities within 4 binomial
    standard deviations - overall and on the pairs with 0.05 < p < 0.95 alone, where a reversed comparison
    (p < u instead of p > u) or a wrong pairing would show."""
    temps = [336.0, 338.0, 340.0, 342.0]
    L = len(temps)
    opts = make_options("snodin_unbound.json", simulation_type="ut_parallel_tempering", num_reps=L, temps=temps, chem_pot_mults=[1] * L,
                        bias_mults=[1] * L, stacking_mults=[1] * L, exchange_interval=100, swaps=rounds, random_seed=4711)
    sim = Simulation(write_inp(str(tmp_path / "freq.inp"), opts), n_ladders * L, 0, lib=lib)
    eng = sim.engine
    red = np.zeros(64)
    nst = eng.L.ldo_get_reduced_staple_u(eng.h, red.ctypes.data)
    q2r_prev = np.tile(np.arange(L, dtype=np.int32), (n_ladders, 1))
    sum_p = var = n_acc = 0.0
    part_p = part_var = part_acc = 0.0
    w_stat = w_var = w_power = 0.0  # sum (swapped - p)(2p - 1): a reversed comparison shifts it by -sum (2p - 1)^2
    n_part = 0
    for swap_i in range(1, rounds + 1):
        assert sim.exchange_advance() == 0
        dep = np.ascontiguousarray(eng.exchange_collect().reshape(n_ladders, L, -1))
        sim.exchange_apply(swap_i)
        q2r = sim.exchange_state(n_ladders, L)[0]
        for l in range(n_ladders):
            for i in range(swap_i % 2, L - 1, 2):
                r1, r2 = q2r_prev[l, i], q2r_prev[l, i + 1]
                d1, d2 = dep[l, r1], dep[l, r2]
                p = eng.L.ldo_exchange_acceptance_p(nst, red.ctypes.data, temps[i], temps[i + 1], d1[3], d2[3], 1.0, 1.0,
                                                    d1.ctypes.data, d2.ctypes.data)
                swapped = q2r[l, i] == r2 and q2r[l, i + 1] == r1
                assert swapped or (q2r[l, i] == r1 and q2r[l, i + 1] == r2)
                sum_p += p
                var += p * (1 - p)
                n_acc += swapped
                if 0.05 < p < 0.95:
                    n_part += 1
                    part_p += p
                    part_var += p * (1 - p)
                    part_acc += swapped
                    w_stat += (float(swapped) - p) * (2 * p - 1)
                    w_var += p * (1 - p) * (2 * p - 1) ** 2
                    w_power += (2 * p - 1) ** 2
        q2r_prev = q2r.copy()
    eng.assert_ok()
    assert n_part >= 30, n_part
    assert abs(n_acc - sum_p) <= 4 * np.sqrt(var), (n_acc, sum_p, var)
    assert abs(part_acc - part_p) <= 4 * np.sqrt(part_var), (part_acc, part_p, part_var, n_part)
    # a reversed comparison (accepting with 1 - p) would move the weighted statistic by -w_power: the test can tell
    assert abs(w_stat) <= 4 * np.sqrt(w_var), (w_stat, w_var)
    assert w_power > 8 * np.sqrt(w_var), (w_power, w_var)


def test_swap_acceptance_frequency(hostsim_lib, tmp_path):
    _swap_acceptance_frequency(hostsim_lib, tmp_path, n_ladders=24, rounds=10)


@pytest.mark.gpu
def test_swap_acceptance_frequency_gpu(tmp_path):
    _swap_acceptance_frequency(None, tmp_path, n_ladders=512, rounds=16)


def _stream_ordered_round_matches_stepwise(lib, tmp_path, two_d):
    """ldo_sim_exchange_round (everything enqueued on the engine's stream, map and counters resident on the device)
    takes the decisions and trajectories of the advance / collect / apply sequence the tests above pin."""
    if two_d:
        opts = make_options("snodin_unbound.json", simulation_type="2d_parallel_tempering", num_reps=4, temps=[332.0, 338.0],
                            stacking_mults=[1.0, 0.8], exchange_interval=30, swaps=4, random_seed=5)
        L = 4
    else:
        opts, L = pt_options(), len(TEMPS)
    inp = write_inp(str(tmp_path / "r.inp"), opts)
    n_ladders = 7
    a = Simulation(inp, n_ladders * L, 0, lib=lib)
    b = Simulation(inp, n_ladders * L, 0, lib=lib)
    for swap_i in range(1, 8):
        assert a.exchange_advance() == 0
        a.exchange_apply(swap_i)
        assert b.exchange_round(swap_i) == 0
        if swap_i == 4:  # reading the map back in the middle must not disturb the resident copy
            assert np.array_equal(a.exchange_state(n_ladders, L, two_d=two_d)[0], b.exchange_state(n_ladders, L, two_d=two_d)[0])
    for x, y in zip(a.exchange_state(n_ladders, L, two_d=two_d), b.exchange_state(n_ladders, L, two_d=two_d)):
        assert np.array_equal(x, y)
    assert a.exchange_state(n_ladders, L, two_d=two_d)[2].sum() > 0
    assert np.array_equal(a.engine.energies(), b.engine.energies())
    assert np.array_equal(a.engine.control()["temp_idx"], b.engine.control()["temp_idx"])
    # switching back to the stepwise calls continues from the resident state
    assert a.exchange_advance() == 0 and b.exchange_advance() == 0
    a.exchange_apply(8)
    b.exchange_apply(8)
    assert np.array_equal(a.exchange_state(n_ladders, L, two_d=two_d)[0], b.exchange_state(n_ladders, L, two_d=two_d)[0])
    assert np.array_equal(a.engine.energies(), b.engine.energies())


@pytest.mark.parametrize("two_d", [False, True])
def test_stream_ordered_round_matches_stepwise(hostsim_lib, tmp_path, two_d):
    _stream_ordered_round_matches_stepwise(hostsim_lib, tmp_path, two_d)


@pytest.mark.gpu
def test_stream_ordered_round_matches_stepwise_gpu(tmp_path):
    _stream_ordered_round_matches_stepwise(None, tmp_path, False)
