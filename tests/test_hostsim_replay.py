"""Device logic (the .cuh sources of the CUDA kernels) compiled for the host with one emulated lane and
replayed against tapes recorded from the unmodified reference: lattice / domain / staple state bit-exact
after every chunk, energies to 1e-12. This is a CPU check of the *sources*; the GPU tests
(tests/test_gpu_parity.py) run the same fixtures through the CUDA library."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, INPUTS, make_options, options_from_fixture, replay_fixture_through, write_inp, assert_state_equal, live_replay
from latticednaorigami_b200.binding import Simulation


@pytest.mark.parametrize("name", ["four_unbound_340K", "snodin_assembled_330K", "snodin_unbound_335K",
                                  "snodin_assembled_ctcb_332K", "snodin_unbound_ctcb_334K",
                                  "snodin_assembled_linker_341K", "snodin_unbound_linker_336K"])
def test_replay_fixture(hostsim_lib, tmp_path, name):
    fx = np.load(os.path.join(GOLDEN, f"replay_{name}.npz"))
    inp = write_inp(str(tmp_path / "r.inp"), options_from_fixture(fx))
    sim = Simulation(inp, 2, 0, lib=hostsim_lib)
    replay_fixture_through(sim, fx, replicas=[0, 1])


def test_live_replay_against_oracle(hostsim_lib, oracle, tmp_path):
    """Fresh seeds each run: record from the reference, replay through the device logic."""
    for system, temp, seed, steps, moveset in [("snodin_assembled.json", 330, 101, 60, "moveset_standard.json"),
                                               ("snodin_unbound.json", 340, 102, 400, "moveset_standard.json"),
                                               ("snodin_assembled.json", 338, 103, 400, "moveset_ctcb.json")]:
        opts = make_options(system, moveset, temp=temp)
        r = oracle.RefSystem(opts)
        r.seed(seed)
        sim = Simulation(write_inp(str(tmp_path / f"{seed}.inp"), opts), 1, 0, lib=hostsim_lib)
        for _ in range(4):
            r.tape(clear=True)
            r.simulate(steps // 4)
            tape = r.tape(clear=True)
            sim.engine.attach_tape(0, tape)
            sim.engine.run(steps // 4, 0, 0, 0)
            sim.engine.assert_ok()
            assert sim.engine.tape_position(0) == len(tape)
            assert_state_equal(sim.engine.state(0), r.state(), f"{system} seed {seed}")
            e = r.energy()
            assert abs(sim.engine.energies()[0, 0] - e) <= 1e-12 * max(1.0, abs(e))


def test_centering_and_constraint_check(hostsim_lib, oracle, tmp_path):
    opts = make_options("snodin_assembled.json", temp=330, centering_freq=7, constraint_check_freq=5)
    r = oracle.RefSystem(opts)
    r.seed(9)
    r.simulate(40)
    tape = r.tape()
    sim = Simulation(write_inp(str(tmp_path / "c.inp"), opts), 1, 0, lib=hostsim_lib)
    sim.engine.attach_tape(0, tape)
    sim.engine.run(40, 7, 0, 5)
    sim.engine.assert_ok()
    assert_state_equal(sim.engine.state(0), r.state())
    assert abs(sim.engine.energies()[0, 0] - r.energy()) <= 1e-12 * abs(r.energy())


def test_tape_mismatch_is_detected(hostsim_lib, tmp_path):
    fx = np.load(os.path.join(GOLDEN, "replay_four_unbound_340K.npz"))
    sim = Simulation(write_inp(str(tmp_path / "m.inp"), options_from_fixture(fx)), 1, 0, lib=hostsim_lib)
    tape = fx["tape"][: int(fx["tape_lens"][0])].copy()
    tape["hi"][np.nonzero(tape["kind"] == 1)[0][3]] += 1
    sim.engine.attach_tape(0, tape)
    sim.engine.run(int(fx["chunk"]))
    st, _ = sim.engine.status()
    assert st[0] == 2


def distance_order_params_and_biases(lib, oracle, tmp_path):
    """Dist / AdjacentSite / Sum order parameters of the move-update kind with well biases on them
    (order_params.cpp:34-161, bias_functions.cpp:114-222): replay against the live oracle; the biases enter every
    acceptance test, so the lattice state stays bit-exact only if they match, and the order-parameter values
    and the total bias are compared directly as well."""
    tags = ["numstaples", "numfulldomains", "dist-ends", "dist-mid", "adj-2-9", "dist-sum"]
    seen = set()
    for system, temp, seed in [("snodin_unbound.json", 338, 5), ("snodin_assembled.json", 341, 6)]:
        opts = make_options(system, temp=temp, bias_functions_file=os.path.join(GOLDEN, "inputs", "biases_dist.json"))
        opts["order_parameter_file"] = os.path.join(GOLDEN, "inputs", "ops_dist.json")
        r = oracle.RefSystem(opts)
        r.seed(seed)
        sim = Simulation(write_inp(str(tmp_path / f"d{seed}.inp"), opts), 1, 0, lib=lib)
        for k in range(24):
            r.tape(clear=True)
            r.simulate(25)
            tape = r.tape(clear=True)
            sim.engine.attach_tape(0, tape)
            sim.engine.run(25, 0, 0, 0)
            sim.engine.assert_ok()
            assert sim.engine.tape_position(0) == len(tape)
            assert_state_equal(sim.engine.state(0), r.state(), f"{system} chunk {k}")
            got = sim.engine.order_params()[0]
            for i, tag in enumerate(tags):
                assert got[i] == r.order_param(tag), (tag, got)
            assert abs(sim.engine.energies()[0, 4] - r.total_bias()) < 1e-12
            seen.add((int(got[2]), int(got[3]), int(got[4]), round(r.total_bias(), 6)))
    assert len(seen) > 8  # the distances and the biases actually move


def test_distance_order_params_and_biases(hostsim_lib, oracle, tmp_path):
    distance_order_params_and_biases(hostsim_lib, oracle, tmp_path)


@pytest.mark.gpu
def test_distance_order_params_and_biases_gpu(oracle, tmp_path):
    """GPU twin (row a24): the same live-oracle comparison through the CUDA library."""
    distance_order_params_and_biases(None, oracle, tmp_path)


# ---- potential options no shipped input uses: Disallowed misbinding, the mean-field correction -----------------

def disallowed_misbinding(lib, oracle, tmp_path):
    """misbinding_pot=Disallowed (DisallowedMisbindingPotential, origami_potential.cpp:947-950): every misbound
    placement violates; replayed against the live oracle from both snodin starts."""
    for system, temp, seed, steps in [("snodin_unbound.json", 337, 61, 1200), ("snodin_assembled.json", 339, 62, 400)]:
        r, sim = live_replay(oracle, tmp_path, lib, make_options(system, temp=temp, misbinding_pot="Disallowed"), seed, steps, name="dis")
        assert r.counters()["misbound_pairs"] == 0
    # the option matters: with Opposing the same seed misbinds along the way
    r = oracle.RefSystem(make_options("snodin_unbound.json", temp=337))
    r.seed(61)
    seen = 0
    for _ in range(12):
        r.simulate(100)
        seen = max(seen, r.counters()["misbound_pairs"])
    assert seen > 0


def mean_field_correction(lib, oracle, tmp_path):
    """apply_mean_field_cor=true (origami_system.cpp:387-400, 858-868; origami_potential.cpp:1060-1100): the log 6 terms of
    chain insertion and of the first two fully bound pairs enter energies and acceptance ratios."""
    for system, temp, seed, steps in [("snodin_unbound.json", 336, 71, 1600), ("four_unbound.json", 338, 72, 2400)]:
        kw = dict(temp=temp, apply_mean_field_cor=True)
        if system == "four_unbound.json":
            opts = make_options(system, "moveset_four.json", max_total_staples=2, max_type_staples=2, **kw)
        else:
            opts = make_options(system, **kw)
        r, sim = live_replay(oracle, tmp_path, lib, opts, seed, steps, chunks=8, name="mf")
        split = r.energy_split()
        e = sim.engine.energies()[0]
        assert abs(e[1] - split["enthalpy"]) <= 1e-12 * max(1.0, abs(split["enthalpy"]))
        assert abs(e[2] - split["entropy"]) <= 1e-12 * max(1.0, abs(split["entropy"]))
    assert r.counters()["staples"] >= 1  # staples were inserted: the chain terms were exercised


def adaptive_exchange(lib, oracle, tmp_path):
    """adaptive_exchange = true (met_movetypes.cpp:228-234, 275-282): an exchange multiplier that makes an acceptance
    probability exceed one is divided by ten and the move rejected - state of the movetype, kept per replica on the
    device. Replayed against the live oracle with multipliers large enough to be cut along the way."""
    import json
    ms = json.load(open(os.path.join(INPUTS, "moveset_standard.json")))
    for mt in ms["origami"]["movetypes"]:
        if mt["type"] == "MetStapleExchange":
            mt["adaptive_exchange"] = True
            mt["exchange_mults"] = [400.0] * 12
            mt["freq"] = "3/8"
        else:
            mt["freq"] = {"OrientationRotation": "2/8"}.get(mt["type"], "1/8")
    path = str(tmp_path / "ms_adaptive.json")
    with open(path, "w") as f:
        json.dump(ms, f)
    cut = 0
    for system, temp, seed, steps in [("snodin_unbound.json", 331, 81, 2000), ("snodin_assembled.json", 350, 82, 1200)]:
        opts = make_options(system, temp=temp)
        opts["movetype_file"] = path
        r, sim = live_replay(oracle, tmp_path, lib, opts, seed, steps, name="adapt")
        mults = sim.engine.exchange_mults(1)
        assert mults.shape == (1, 12) and set(mults[0]) <= {400.0, 40.0, 4.0, 0.4}
        cut += int((mults[0] != 400.0).sum())
    assert cut > 0, "no multiplier was adapted: the test does not exercise the branch"


def test_adaptive_exchange_multipliers(hostsim_lib, oracle, tmp_path):
    adaptive_exchange(hostsim_lib, oracle, tmp_path)


def test_disallowed_misbinding(hostsim_lib, oracle, tmp_path):
    disallowed_misbinding(hostsim_lib, oracle, tmp_path)


def test_mean_field_correction(hostsim_lib, oracle, tmp_path):
    mean_field_correction(hostsim_lib, oracle, tmp_path)


@pytest.mark.gpu
def test_disallowed_misbinding_and_mean_field_gpu(oracle, tmp_path):
    disallowed_misbinding(None, oracle, tmp_path)
    mean_field_correction(None, oracle, tmp_path)
    adaptive_exchange(None, oracle, tmp_path)
