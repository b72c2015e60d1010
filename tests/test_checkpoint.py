"""Checkpoint blobs (ldo_checkpoint_save / ldo_checkpoint_load): one self-contained blob per replica, so any
sub-range of a saved buffer restores the replicas it names - into any replica index, of another engine too."""
import numpy as np
import pytest

from conftest import INPUTS, assert_state_equal, make_options, write_inp
from latticednaorigami_b200.binding import Simulation
import os


def _partial_range_roundtrip(tmp_path, lib):
    opts = make_options("snodin_assembled.json", temp=334, random_seed=5)
    inp = write_inp(str(tmp_path / "k.inp"), opts)
    a = Simulation(inp, 6, 0, lib=lib)
    a.engine.run(60)
    size = a.engine.checkpoint_size()
    blob = a.engine.checkpoint_save()
    assert blob.size == 6 * size
    a.engine.run(40)
    # replicas 2..3 of the saved buffer go to replicas 0..1 of a fresh engine, replica 5 to replica 3
    b = Simulation(inp, 4, 0, lib=lib)
    b.engine.checkpoint_load(blob[2 * size:4 * size], first=0, count=2)
    b.engine.checkpoint_load(blob[5 * size:6 * size], first=3, count=1)
    b.engine.synchronize()
    b.engine.run(40)
    b.engine.assert_ok()
    ea, eb = a.engine.energies(), b.engine.energies()
    for src, dst in ((2, 0), (3, 1), (5, 3)):
        assert np.array_equal(ea[src], eb[dst])
        assert_state_equal(b.engine.state(dst), a.engine.state(src), f"{src}->{dst}")
    # a sub-range save equals the corresponding slice of the full save
    part = a.engine.checkpoint_save(first=1, count=3)
    full = a.engine.checkpoint_save()
    assert np.array_equal(part, full[size:4 * size])


def _tape_is_not_part_of_a_checkpoint(tmp_path, lib):
    fx = np.load(os.path.join(os.path.dirname(INPUTS), "replay_snodin_assembled_330K.npz"))
    from conftest import options_from_fixture
    a = Simulation(write_inp(str(tmp_path / "t.inp"), options_from_fixture(fx)), 2, 0, lib=lib)
    a.engine.attach_tape(0, fx["tape"][:int(fx["tape_lens"][0])])
    blob = a.engine.checkpoint_save()
    a.engine.checkpoint_load(blob)
    a.engine.synchronize()
    assert a.engine.tape_position(0) == 0
    a.engine.seed(3)
    a.engine.run(5)  # would dereference a dangling tape pointer if one had survived the round trip
    a.engine.assert_ok()


def _grid_state_travels_with_the_blob(tmp_path, lib):
    from test_umbrella_sampling import us_options
    opts = us_options(tmp_path, "mw_umbrella_sampling", output_filebase="")
    inp = write_inp(str(tmp_path / "g.inp"), opts)
    a = Simulation(inp, 3, 0, lib=lib)
    n = 37  # box of numfulldomains the host driver allocates
    vals = np.full(n, np.nan)
    vals[:5] = [0.5, -1.0, 2.0, 0.25, -0.75]
    a.engine.set_grid_bias(1, 1, [0], [n], vals)
    a.engine.run(30)
    visits = a.engine.grid_visits(1, 1, n)
    assert visits.sum() == 30
    size = a.engine.checkpoint_size()
    blob = a.engine.checkpoint_save()
    b = Simulation(inp, 3, 0, lib=lib)
    b.engine.checkpoint_load(blob[size:2 * size], first=2, count=1)
    b.engine.synchronize()
    assert np.array_equal(b.engine.grid_visits(2, 1, n), visits)
    a.engine.run(20)
    b.engine.run(20)
    assert np.array_equal(a.engine.energies()[1], b.engine.energies()[2])  # same grid values and window -> same trajectory
    assert np.array_equal(a.engine.grid_visits(1, 1, n), b.engine.grid_visits(2, 1, n))


def test_partial_range_roundtrip(hostsim_lib, tmp_path):
    _partial_range_roundtrip(tmp_path, hostsim_lib)


def test_tape_is_not_part_of_a_checkpoint(hostsim_lib, tmp_path):
    _tape_is_not_part_of_a_checkpoint(tmp_path, hostsim_lib)


def test_grid_state_travels_with_the_blob(hostsim_lib, tmp_path):
    _grid_state_travels_with_the_blob(tmp_path, hostsim_lib)


@pytest.mark.gpu
def test_checkpoint_blobs_gpu(tmp_path):
    _partial_range_roundtrip(tmp_path, None)
    _tape_is_not_part_of_a_checkpoint(tmp_path, None)
    _grid_state_travels_with_the_blob(tmp_path, None)
