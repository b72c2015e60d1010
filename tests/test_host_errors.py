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


def test_umbrella_sampling_restart_inputs_that_are_refused(hostsim_lib, tmp_path):
    """read_biases names a file that must exist (us_simulation.cpp:52-57); restart_us_iter needs Boost archives."""
    base = dict(temp=334, simulation_type="umbrella_sampling", bias_functions_file=os.path.join(INPUTS, "biases_mwus-numfulldomains.json"),
                bias_functions_mult=1, us_grid_bias_tag="grid", max_num_iters=1, max_D_bias=1.5, equil_steps=10, max_equil_dur=1000,
                iter_steps=10, max_iter_dur=1000, output_filebase=str(tmp_path / "us"))
    sim = Simulation(write_inp(str(tmp_path / "u1.inp"), make_options("snodin_unbound.json", read_biases=True,
                                                                       biases_file=str(tmp_path / "nothing.biases"), **base)), 1, 0, lib=hostsim_lib)
    with pytest.raises(LdoError, match="Restart bias file .* does not exist"):
        sim.run()
    sim = Simulation(write_inp(str(tmp_path / "u2.inp"), make_options("snodin_unbound.json", restart_us_iter=True,
                                                                       restart_us_filebase=str(tmp_path / "x"), **base)), 1, 0, lib=hostsim_lib)
    with pytest.raises(LdoError, match="restart_us_iter"):
        sim.run()


def test_move_trackers_are_optional(hostsim_lib, tmp_path):
    """Without output files the typed trackers are off (no extra per-replica memory); asking for them then is an error."""
    sim = Simulation(write_inp(str(tmp_path / "t.inp"), make_options("snodin_unbound.json", temp=340)), 2, 0, lib=hostsim_lib)
    sim.engine.run(50)
    with pytest.raises(LdoError, match="not enabled"):
        sim.engine.move_trackers(0)
    att0, _ = sim.engine.move_stats()
    sim.engine.enable_move_trackers(True)
    sim.engine.run(200)
    sticky, counts = sim.engine.move_trackers(1)
    att1, acc1 = sim.engine.move_stats()
    # every attempt made since the trackers were switched on is in exactly one bin of its movetype (standard moveset:
    # one field each); the orientation rotation has no tracker
    assert counts[0].sum() == 0
    for i in range(1, 5):
        assert counts[i, :, :, 0].sum() == att1[1][i] - att0[1][i]
    assert counts[1:, :, :, 0].sum() > 0
    sim.engine.enable_move_trackers(False)
    with pytest.raises(LdoError, match="not enabled"):
        sim.engine.move_trackers(0)


def test_wall_clock_limits_stop_the_drivers(hostsim_lib, tmp_path, capfd):
    """max_duration (simulation.cpp:574-578: checked after every step, "Maximum time allowed reached") and max_pt_dur
    (ptmc_simulation.cpp:116-128) end a run early and cleanly: the files written so far are complete and the summary is
    written."""
    import time
    opts = make_options("snodin_unbound.json", temp=340, random_seed=3, ct_steps=10 ** 9, max_duration=0.2, configs_output_freq=50,
                        output_filebase=str(tmp_path / "ct"))
    sim = Simulation(write_inp(str(tmp_path / "ct.inp"), opts), 1, 0, lib=hostsim_lib)
    t0 = time.time()
    sim.run()
    assert time.time() - t0 < 20 and 0 < sim.step < 10 ** 9 and sim.step % 50 == 0
    frames = [f for f in (tmp_path / "ct.trj").read_text().split("\n\n") if f.strip()]
    # the limit is tested before the outputs of the step it stops at (simulation.cpp:621-625, 640-646)
    assert len(frames) == sim.step // 50 - 1 and (tmp_path / "ct.moves").exists()
    assert "Maximum time allowed reached" in capfd.readouterr().out

    opts = make_options("snodin_unbound.json", simulation_type="ut_parallel_tempering", random_seed=3, temps=[330, 335, 340], num_reps=3,
                        chem_pot_mults=[1, 1, 1], bias_mults=[1, 1, 1], stacking_mults=[1, 1, 1], exchange_interval=20, swaps=10 ** 8,
                        max_pt_dur=0.2, configs_output_freq=20, output_filebase=str(tmp_path / "pt"))
    sim = Simulation(write_inp(str(tmp_path / "pt.inp"), opts), 3, 0, lib=hostsim_lib)
    t0 = time.time()
    sim.run()
    assert time.time() - t0 < 20 and 0 < sim.step < 20 * 10 ** 8 and sim.step % 20 == 0
    swp = (tmp_path / "pt.swp").read_text().strip().splitlines()
    assert len(swp) >= 2 and all(sorted(int(x) for x in row.split()) == [0, 1, 2] for row in swp[1:])
    assert "Maximum time allowed reached" in capfd.readouterr().out


def tracked_kernels_same_trajectory(lib, tmp_path, replicas, large=False):
    """The Tracked<K> instantiation of the kernels (launched while the typed trackers are on) and the production one are
    the same move code: a run that switches between them is the run that never does (Philox draws, same seed). `large`:
    the 168-domain raster, i.e. the in-place kernel."""
    import numpy as np
    from conftest import assert_state_equal
    opts = make_options("snodin_assembled.json", "moveset_linker.json", temp=338, random_seed=17)
    if large:
        from synthetic import UNIFORM_OPTIONS, write_raster_system
        # cold and concentrated enough that staples bind within the run
        opts = make_options(temp=270, max_total_staples=168, max_type_staples=2, staple_M=1.0, random_seed=17, **UNIFORM_OPTIONS)
        opts["origami_input_filename"] = write_raster_system(str(tmp_path / "raster.json"), 12, 14, False)
    inp = write_inp(str(tmp_path / "k.inp"), opts)
    a = Simulation(inp, replicas, 0, lib=lib)
    b = Simulation(inp, replicas, 0, lib=lib)
    a.engine.run(600)
    b.engine.run(150)
    b.engine.enable_move_trackers(True)
    b.engine.run(300)
    b.engine.enable_move_trackers(False)
    b.engine.run(150)
    a.engine.assert_ok()
    b.engine.assert_ok()
    for r in range(replicas):
        assert_state_equal(b.engine.state(r), a.engine.state(r), f"replica {r}")
    ea, eb = a.engine.energies(), b.engine.energies()
    assert np.all(np.abs(ea - eb) <= 1e-12 * np.maximum(1.0, np.abs(ea)))
    assert np.array_equal(a.engine.rng_state(), b.engine.rng_state())
    assert all(np.array_equal(x, y) for x, y in zip(a.engine.move_stats(), b.engine.move_stats()))
    assert len(a.engine.state(0)["chain_len"]) > 1  # staples are bound: the run is not a free scaffold only


def test_tracked_kernels_same_trajectory(hostsim_lib, tmp_path):
    tracked_kernels_same_trajectory(hostsim_lib, tmp_path, 2)


def test_tracked_kernels_same_trajectory_in_place(hostsim_lib, tmp_path):
    tracked_kernels_same_trajectory(hostsim_lib, tmp_path, 1, large=True)


@pytest.mark.gpu
def test_tracked_kernels_same_trajectory_gpu(tmp_path):
    tracked_kernels_same_trajectory(None, tmp_path, 64)
    (tmp_path / "large").mkdir()
    tracked_kernels_same_trajectory(None, tmp_path / "large", 16, large=True)
