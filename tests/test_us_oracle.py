"""Adaptive umbrella sampling against the reference's OWN drivers (row a27): SimpleUSGCMCSimulation and
MWUSGCMCSimulation (us_simulation.cpp:75-170, 262-305, 385-452, 475-553) run in the oracle - the multi-window
driver with one thread per window over the thread-backed boost::mpi shim - with every window's draws taped. The host
driver replays the tapes on the engine and must write the same bias files after every iteration
(estimate_current_weights + update_bias, read back as JSON: the reference prints entries in unordered_map order and
with std::to_string's 6 decimals), the same per-window summaries, and end in the same lattice states.
CPU: host emulation; GPU: the CUDA library."""
import json
import os

import numpy as np
import pytest

from conftest import INPUTS, make_options, write_inp
from latticednaorigami_b200.binding import Simulation


def us_options(tmp_path, base, sim_type, **kw):
    opts = make_options("snodin_unbound.json", temp=334, simulation_type=sim_type,
                        bias_functions_file=os.path.join(INPUTS, "biases_mwus-numfulldomains.json"), bias_functions_mult=1,
                        us_grid_bias_tag="grid", max_num_iters=3, max_D_bias=1.5, equil_steps=600, max_equil_dur=100000,
                        iter_steps=1200, max_iter_dur=100000, output_filebase=str(tmp_path / base))
    if sim_type != "umbrella_sampling":
        opts.update(multi_window=True, windows_file=os.path.join(INPUTS, "snodin-numfulldomains.windows"))
    opts.update(kw)
    return opts


def read_biases(path):
    return {tuple(b["point"]): b["bias"] for b in json.load(open(path))["biases"]}


def us_against_oracle(oracle, tmp_path, lib, sim_type):
    n = 1 if sim_type == "umbrella_sampling" else 3
    ref_opts = us_options(tmp_path, "ref", sim_type)
    ref = oracle.us_run(ref_opts, n, [300 + 11 * r for r in range(n)], workdir=str(tmp_path))
    sim = Simulation(write_inp(str(tmp_path / "our.inp"), us_options(tmp_path, "our", sim_type, random_seed=1)), n, 0, lib=lib)
    for r in range(n):
        sim.engine.attach_tape(r, ref["tapes"][r])
    sim.run()
    sim.engine.assert_ok()
    posts = [""] if n == 1 else ["_win-0--4", "_win-2--6", "_win-4--8"]
    visited = 0
    for r, post in enumerate(posts):
        assert sim.engine.tape_position(r) == len(ref["tapes"][r]), (sim_type, r)
        got = sim.engine.state(r)
        for k in ("chain_index", "chain_ident", "chain_len", "pos", "ore"):
            assert np.array_equal(got[k], ref["states"][r][k]), (sim_type, r, k)
        for it in range(3):
            for tail in (f"_iter-{it}-inp.biases", f"_iter-{it}.biases"):
                want = read_biases(tmp_path / f"ref{post}{tail}")
                have = read_biases(tmp_path / f"our{post}{tail}")
                assert want.keys() == have.keys(), (sim_type, post, tail)
                for pt in want:
                    assert abs(want[pt] - have[pt]) <= 1e-6, (sim_type, post, tail, pt, want[pt], have[pt])
                visited = max(visited, len(want))
        # the per-window summary stream ("Iteration: n / Gridpoint w, P, E", us_simulation.cpp:436-452)
        ref_out = (tmp_path / f"ref{post}.out").read_text() if n > 1 else None
        if ref_out is not None:
            our_out = (tmp_path / f"our{post}.out").read_text()
            assert our_out == ref_out, (sim_type, post)
    assert visited >= 3  # several grid points were visited and re-weighted


@pytest.mark.parametrize("sim_type", ["umbrella_sampling", "mw_umbrella_sampling"])
def test_bias_update_matches_reference_driver(hostsim_lib, oracle, tmp_path, sim_type):
    us_against_oracle(oracle, tmp_path, hostsim_lib, sim_type)


@pytest.mark.gpu
@pytest.mark.parametrize("sim_type", ["umbrella_sampling", "mw_umbrella_sampling"])
def test_bias_update_matches_reference_driver_gpu(oracle, tmp_path, sim_type):
    us_against_oracle(oracle, tmp_path, None, sim_type)
