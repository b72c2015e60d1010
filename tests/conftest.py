import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
INPUTS = os.path.join(GOLDEN, "inputs")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

HOSTSIM_LIB = os.path.join(ROOT, "tests", "hostsim", "libldo_hostsim.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: minutes of GPU time (long equilibrium runs)")


# Options of examples/constant-temp.inp with every output off (same dict as oracle_ref.SNODIN_OPTIONS)
BASE_OPTIONS = {
    "domain_type": "HalfTurn", "binding_pot": "FourBody", "misbinding_pot": "Opposing", "stacking_pot": "Constant",
    "hybridization_pot": "NearestNeighbour", "apply_mean_field_cor": False, "temp": 330, "staple_M": 1e-7,
    "cation_M": 0.5, "staple_u_mult": 1, "stacking_ene": -1000, "max_total_staples": 24, "max_type_staples": 12,
    "max_staple_size": 2, "domain_update_biases_present": False, "simulation_type": "constant_temp",
    "centering_freq": 0, "constraint_check_freq": 0, "max_duration": 1e9, "ct_steps": 0, "logging_freq": 0,
    "configs_output_freq": 0, "vtf_output_freq": 0, "vcf_per_domain": False, "counts_output_freq": 0,
    "order_params_output_freq": 0, "times_output_freq": 0, "energies_output_freq": 0,
}


def make_options(system="snodin_unbound.json", moveset="moveset_standard.json", **kw):
    d = dict(BASE_OPTIONS)
    d["origami_input_filename"] = os.path.join(INPUTS, system)
    d["order_parameter_file"] = os.path.join(INPUTS, "ops_standard.json")
    d["movetype_file"] = os.path.join(INPUTS, moveset)
    d.update(kw)
    return d


def write_inp(path, options):
    with open(path, "w") as f:
        for k, v in options.items():
            if isinstance(v, bool):
                v = "true" if v else "false"
            elif isinstance(v, (list, tuple)):
                v = " ".join(str(x) for x in v)
            f.write(f"{k}={v}\n")
    return path


def options_from_fixture(fx):
    """Options stored in a replay fixture (file names are relative to tests/golden/inputs)."""
    opts = json.loads(str(fx["options"]))
    for k in ("origami_input_filename", "order_parameter_file", "movetype_file"):
        opts[k] = os.path.join(INPUTS, opts[k])
    opts.pop("output_filebase", None)
    return opts


def fixture_state(fx, i):
    nc, nd = int(fx["n_chains"][i]), int(fx["n_domains"][i])
    return {"chain_index": fx["chain_index"][i][:nc], "chain_ident": fx["chain_ident"][i][:nc],
            "chain_len": fx["chain_len"][i][:nc], "pos": fx["pos"][i][:nd], "ore": fx["ore"][i][:nd],
            "state": fx["state"][i][:nd], "bound": fx["bound"][i][:nd]}


def assert_state_equal(got, want, where=""):
    for k in ("chain_index", "chain_ident", "chain_len", "pos", "ore", "state", "bound"):
        assert got[k].shape == want[k].shape, f"{where}: {k} shape {got[k].shape} != {want[k].shape}"
        assert np.array_equal(got[k], want[k]), f"{where}: {k} differs"


def replay_fixture_through(sim, fx, replica=0, replicas=None):
    """Feed a recorded tape chunk by chunk and compare the lattice state after every chunk (bit-exact)
    and the running energy (1e-12 relative, north_star tolerance)."""
    eng = sim.engine
    replicas = [replica] if replicas is None else replicas
    tape = fx["tape"]
    off = 0
    chunk = int(fx["chunk"])
    for i, n in enumerate(fx["tape_lens"]):
        part = tape[off:off + int(n)]
        off += int(n)
        for r in replicas:
            eng.attach_tape(r, part)
        eng.run(chunk)
        eng.assert_ok()
        want = fixture_state(fx, i)
        energies = eng.energies()
        counters = eng.counters()
        for r in replicas:
            assert eng.tape_position(r) == int(n), f"chunk {i}: tape not fully consumed"
            assert_state_equal(eng.state(r), want, f"chunk {i} replica {r}")
            e_ref = float(fx["energy"][i])
            assert abs(energies[r, 0] - e_ref) <= 1e-12 * max(1.0, abs(e_ref)), f"chunk {i}: energy {energies[r, 0]} vs {e_ref}"
            assert list(counters[r]) == list(fx["counters"][i]), f"chunk {i}: counters"
    att, acc = eng.move_stats()
    for r in replicas:
        assert list(att[r]) == list(fx["attempts"]) and list(acc[r]) == list(fx["accepts"])


_hostsim = None


def load_hostsim():
    """Host emulation of the device sources (one emulated lane): test infrastructure, built on demand and
    opened HERE - the package's own loader refuses anything but a CUDA build. Returns a bound library
    handle for Simulation(..., lib=...)."""
    global _hostsim
    if _hostsim is None:
        import ctypes

        from latticednaorigami_b200 import binding
        csrc = os.path.join(ROOT, "latticednaorigami_b200", "csrc")
        srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
        newest = max(os.path.getmtime(p) for p in srcs)
        if not os.path.exists(HOSTSIM_LIB) or os.path.getmtime(HOSTSIM_LIB) < newest:
            subprocess.run(["make", "-C", ROOT, "hostsim"], check=True, capture_output=True)
        _hostsim = binding.bind(ctypes.CDLL(HOSTSIM_LIB))
        assert _hostsim.ldo_build_info().decode().startswith("hostsim")
    return _hostsim


@pytest.fixture(scope="session")
def hostsim_lib():
    return load_hostsim()


@pytest.fixture(scope="session")
def oracle():
    import oracle_ref
    if not oracle_ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference; see oracle/Makefile)")
    return oracle_ref


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def live_replay(oracle, tmp_path, lib, opts, seed, steps, chunks=4, replicas=1, name="live"):
    """Record `steps` moves of the live oracle (unmodified reference) in `chunks` tapes and replay them through the
    engine: lattice state bit-exact and energy to 1e-12 after every chunk, move statistics at the end.
    Returns (reference system, simulation)."""
    from latticednaorigami_b200.binding import Simulation
    r = oracle.RefSystem(opts)
    r.seed(seed)
    sim = Simulation(write_inp(str(tmp_path / f"{name}{seed}.inp"), opts), replicas, 0, lib=lib)
    scale = 1.0
    for k in range(chunks):
        r.tape(clear=True)
        r.simulate(steps // chunks)
        tape = r.tape(clear=True)
        for rep in range(replicas):
            sim.engine.attach_tape(rep, tape)
        sim.engine.run(steps // chunks)
        sim.engine.assert_ok()
        want = r.state()
        e = r.energy()
        got_e = sim.engine.energies()
        for rep in range(replicas):
            assert sim.engine.tape_position(rep) == len(tape), f"{name} chunk {k}: tape not fully consumed"
            assert_state_equal(sim.engine.state(rep), want, f"{name} seed {seed} chunk {k}")
            # 1e-12 relative to the terms the RUNNING sum has been made of since it was last rebuilt (no constraint
            # check in these runs): the energy is enthalpy / T minus entropy, each of which can be orders of magnitude
            # larger than their difference, and both codes carry the rounding of everything added and removed so far
            sp = r.energy_split()
            scale = max(scale, abs(e), abs(sp["enthalpy"]), abs(sp["entropy"]))
            assert abs(got_e[rep, 0] - e) <= 1e-12 * scale, (name, k, got_e[rep, 0], e, scale)
        c = r.counters()
        assert [int(x) for x in sim.engine.counters()[0]] == [c[k2] for k2 in ("staples", "domains", "bound_pairs", "fully_bound_pairs",
                                                                              "self_bound_pairs", "misbound_pairs", "stacked_pairs",
                                                                              "unassigned", "current_c_i")]
    att, acc = sim.engine.move_stats()
    ra, rb = r.move_stats()
    assert list(att[0]) == list(ra) and list(acc[0]) == list(rb)
    return r, sim
