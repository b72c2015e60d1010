"""ThreeQuarterTurn domains, Uniform hybridization and the large-system (in-place, HBM-resident) code path,
replayed against the live oracle. CPU: host emulation of the device sources; GPU: the CUDA library."""
import os

import numpy as np
import pytest

from conftest import INPUTS, assert_state_equal, has_cuda, make_options, write_inp
from latticednaorigami_b200.binding import Simulation
from synthetic import UNIFORM_OPTIONS, write_raster_system


def _options(tmp_path, width, n_rows, temp, max_total, cyclic=False, staple_M=1e-5, **kw):
    system = write_raster_system(str(tmp_path / f"raster_{width}x{n_rows}_{int(cyclic)}.json"), width, n_rows, cyclic)
    opts = make_options(temp=temp, max_total_staples=max_total, max_type_staples=2, staple_M=staple_M, **UNIFORM_OPTIONS)
    opts["origami_input_filename"] = system
    opts.update(kw)
    return opts


def _replay(oracle, opts, tmp_path, lib_path, seed, steps, chunks=4):
    r = oracle.RefSystem(opts)
    r.seed(seed)
    sim = Simulation(write_inp(str(tmp_path / f"syn{seed}.inp"), opts), 2, 0, lib=lib_path)
    for _ in range(chunks):
        r.tape(clear=True)
        r.simulate(steps // chunks)
        tape = r.tape(clear=True)
        for rep in (0, 1):
            sim.engine.attach_tape(rep, tape)
        sim.engine.run(steps // chunks, opts.get("centering_freq", 0), 0, opts.get("constraint_check_freq", 0))
        sim.engine.assert_ok()
        for rep in (0, 1):
            assert sim.engine.tape_position(rep) == len(tape)
            assert_state_equal(sim.engine.state(rep), r.state(), f"seed {seed}")
        e = r.energy()
        assert abs(sim.engine.energies()[0, 0] - e) <= 1e-12 * max(1.0, abs(e))
    return r, sim


def _campaign_cases(oracle, tmp_path, lib):
    """Runs of the replay campaign (tests/stress_replay.py) that a change of the FourBody evaluation once broke: the
    reference's trajectory passes through configurations in which a stacking / steric constraint is violated (its
    unchecked placements apply the terms counted up to the violation), and through moves whose stacking terms cancel up
    to the rounding of its term-by-term sum (which decides whether the acceptance test consumes a draw)."""
    for cyc, moveset, seed in ((False, "moveset_linker.json", 8029), (True, "moveset_linker_heavy.json", 9400)):
        opts = _options(tmp_path, 3, 4, temp=302, max_total=8, cyclic=cyc)
        opts["movetype_file"] = os.path.join(INPUTS, moveset)
        _replay(oracle, opts, tmp_path, lib, seed=seed, steps=600, chunks=6)
    for moveset, temp, seed in (("moveset_standard.json", 348, 3048), ("moveset_linker.json", 342, 4907), ("moveset_linker_heavy.json", 330, 9915)):
        _replay(oracle, make_options("snodin_assembled.json", moveset, temp=temp), tmp_path, lib, seed=seed, steps=300, chunks=6)


def test_campaign_cases_hostsim(hostsim_lib, oracle, tmp_path):
    _campaign_cases(oracle, tmp_path, hostsim_lib)


@pytest.mark.gpu
def test_campaign_cases_gpu(oracle, tmp_path):
    _campaign_cases(oracle, tmp_path, None)


def test_three_quarter_turn_small_hostsim(hostsim_lib, oracle, tmp_path):
    """12-domain ThreeQuarterTurn raster (shared-memory staged path), cold enough that staples bind."""
    opts = _options(tmp_path, 3, 4, temp=300, max_total=8)
    r, sim = _replay(oracle, opts, tmp_path, hostsim_lib, seed=41, steps=1600)
    assert r.counters()["fully_bound_pairs"] > 0  # the run exercises ThreeQuarterTurn binding constraints


def test_cyclic_scaffold_hostsim(hostsim_lib, oracle, tmp_path):
    """Cyclic scaffold: modular chain walk, whole-ring selections, cyclic remaining-step counts."""
    opts = _options(tmp_path, 3, 4, temp=300, max_total=8, cyclic=True)
    r, sim = _replay(oracle, opts, tmp_path, hostsim_lib, seed=42, steps=1600)
    assert r.counters()["fully_bound_pairs"] > 0


def test_large_raster_hostsim(hostsim_lib, oracle, tmp_path):
    """168-domain scaffold, 84 staple types (config 5): large capacities, state kept in place (HBM/L2)."""
    opts = _options(tmp_path, 12, 14, temp=295, max_total=168, staple_M=1e-3, centering_freq=50, constraint_check_freq=40)
    _replay(oracle, opts, tmp_path, hostsim_lib, seed=43, steps=400)


def test_ctcb_moves_on_synthetic_systems_hostsim(hostsim_lib, oracle, tmp_path):
    """CTCB scaffold regrowth (contiguous and jump) on ThreeQuarterTurn, cyclic and large (in-place) systems."""
    ctcb = os.path.join(INPUTS, "moveset_ctcb.json")
    _replay(oracle, _options(tmp_path, 3, 4, temp=300, max_total=8, movetype_file=ctcb), tmp_path, hostsim_lib, seed=51, steps=1200)
    _replay(oracle, _options(tmp_path, 3, 4, temp=300, max_total=8, cyclic=True, movetype_file=ctcb), tmp_path, hostsim_lib,
            seed=52, steps=1200)
    _replay(oracle, _options(tmp_path, 12, 14, temp=295, max_total=168, staple_M=1e-3, movetype_file=ctcb), tmp_path,
            hostsim_lib, seed=53, steps=200)


def _linker_cases(tmp_path):
    """Transform / linker movetypes (CTCBLinkerRegrowth, CTCBClusteredLinkerRegrowth, CTRGLinkerRegrowth) on
    partially assembled rasters, where the clustered selection finds bound runs: ThreeQuarterTurn and
    HalfTurn domains, linear and cyclic scaffolds, both hand-written movesets (one with two recoil levels)."""
    for k, ms in enumerate(["moveset_linker.json", "moveset_linker_heavy.json"]):
        path = os.path.join(INPUTS, ms)
        for cyclic in (False, True):
            yield _options(tmp_path, 3, 4, temp=300, max_total=8, cyclic=cyclic, movetype_file=path), 60 + 4 * k + int(cyclic)
            yield _options(tmp_path, 3, 4, temp=310, max_total=8, cyclic=cyclic, movetype_file=path, domain_type="HalfTurn"), 62 + 4 * k + int(cyclic)


def test_linker_moves_on_synthetic_systems_hostsim(hostsim_lib, oracle, tmp_path):
    accepts = np.zeros(3, dtype=np.int64)
    for opts, seed in _linker_cases(tmp_path):
        r, sim = _replay(oracle, opts, tmp_path, hostsim_lib, seed=seed, steps=800)
        att, acc = sim.engine.move_stats()
        accepts += acc[0][-3:]
    assert np.all(accepts > 0)  # every linker movetype was accepted somewhere: the whole move ran
    _replay(oracle, _options(tmp_path, 12, 14, temp=295, max_total=168, staple_M=1e-3,
                             movetype_file=os.path.join(INPUTS, "moveset_linker.json")), tmp_path, hostsim_lib, seed=71, steps=200)


@pytest.mark.gpu
def test_linker_moves_on_synthetic_systems_gpu(oracle, tmp_path):
    accepts = np.zeros(3, dtype=np.int64)
    for opts, seed in _linker_cases(tmp_path):
        r, sim = _replay(oracle, opts, tmp_path, None, seed=seed + 100, steps=1200)
        att, acc = sim.engine.move_stats()
        accepts += acc[0][-3:]
    assert np.all(accepts > 0)
    # the large (in-place, HBM-resident) instantiation
    _replay(oracle, _options(tmp_path, 12, 14, temp=295, max_total=168, staple_M=1e-3,
                             movetype_file=os.path.join(INPUTS, "moveset_linker.json")), tmp_path, None, seed=171, steps=200)


@pytest.mark.gpu
def test_ctcb_moves_on_synthetic_systems_gpu(oracle, tmp_path):
    ctcb = os.path.join(INPUTS, "moveset_ctcb.json")
    _replay(oracle, _options(tmp_path, 3, 4, temp=300, max_total=8, movetype_file=ctcb), tmp_path, None, seed=54, steps=2000)
    _replay(oracle, _options(tmp_path, 3, 4, temp=300, max_total=8, cyclic=True, movetype_file=ctcb), tmp_path, None,
            seed=55, steps=2000)
    _replay(oracle, _options(tmp_path, 12, 14, temp=295, max_total=168, staple_M=1e-3, movetype_file=ctcb), tmp_path,
            None, seed=56, steps=400)


@pytest.mark.gpu
def test_three_quarter_turn_small_gpu(oracle, tmp_path):
    opts = _options(tmp_path, 3, 4, temp=300, max_total=8)
    _replay(oracle, opts, tmp_path, None, seed=45, steps=2000)
    opts = _options(tmp_path, 3, 4, temp=300, max_total=8, cyclic=True)
    _replay(oracle, opts, tmp_path, None, seed=46, steps=2000)


@pytest.mark.gpu
def test_large_raster_gpu(oracle, tmp_path):
    opts = _options(tmp_path, 12, 14, temp=295, max_total=168, staple_M=1e-3, centering_freq=50, constraint_check_freq=40)
    _replay(oracle, opts, tmp_path, None, seed=47, steps=600)


@pytest.mark.gpu
def test_large_raster_philox_invariants(tmp_path):
    """Config 5 shape on the GPU in production mode: annealing-style cooling with consistent state."""
    opts = _options(tmp_path, 12, 14, temp=330, max_total=168, random_seed=5)
    sim = Simulation(write_inp(str(tmp_path / "big.inp"), opts), 256, 0)
    eng = sim.engine
    eng.run(150, 50, 0, 75)
    eng.assert_ok()
    running = eng.energies()[:, 0]
    recomputed, stacked = eng.recompute_energies()
    assert np.all(np.abs(running - recomputed) <= 1e-9 * np.maximum(1.0, np.abs(recomputed)))
    assert np.array_equal(stacked, eng.counters()[:, 6])
