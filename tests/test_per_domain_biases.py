"""Per-domain-update order parameters and biases: OrigamiSystemWithBias (origami_system.cpp:873-957),
SystemOrderParams::update_one_domain / check_one_domain (order_params.cpp:603-625), SystemBiases::calc_one_domain /
check_one_domain (bias_functions.cpp:489-527) with `domain_update_biases_present=true`, replayed against the live oracle.

What the reference really does with such input (found with the oracle, reproduced here): Dist / AdjacentSite parameters
marked update_per_domain are recomputed whenever one of their domains is unassigned or placed by
set_checked_domain_config - but not by set_domain_config, so after an accepted orientation rotation of a bound pair, a
temperature update or a constraint check they stay UNDEFINED (their biases read 0) until a regrowth move touches the
domain again; and no bias function is ever registered as per-domain (the dependency test reads the absent "bias_funcs"
entry as one tag without domains, bias_functions.cpp:371-384), so candidate checks carry no bias and the wells on these
parameters act at the end of the move like any other. The biases enter every acceptance test, so the lattice stays
bit-exact only if all of that matches; parameter values and the stored total bias are compared directly as well.
Recoil-growth, configurational-bias and linker movesets. CPU: host emulation; GPU: the CUDA library."""
import os

import pytest

from conftest import INPUTS, assert_state_equal, make_options, write_inp
from latticednaorigami_b200.binding import Simulation

TAGS = ["numstaples", "numfulldomains", "dist-mid", "dist-ends", "dist-q", "adj-2-9"]
CASES = [("snodin_unbound.json", "moveset_standard.json", 338, 11, 1500), ("snodin_assembled.json", "moveset_standard.json", 344, 12, 400),
         ("snodin_unbound.json", "moveset_ctcb.json", 336, 13, 1200), ("snodin_assembled.json", "moveset_linker.json", 342, 14, 600)]


def per_domain_biases(lib, oracle, tmp_path, cases=CASES, **extra):
    seen = set()
    undefined_seen = False
    for system, moveset, temp, seed, steps in cases:
        opts = make_options(system, moveset, temp=temp, domain_update_biases_present=True,
                            bias_functions_file=os.path.join(INPUTS, "biases_perdomain.json"), **extra)
        opts["order_parameter_file"] = os.path.join(INPUTS, "ops_perdomain.json")
        r = oracle.RefSystem(opts)
        r.seed(seed)
        sim = Simulation(write_inp(str(tmp_path / f"pd{seed}.inp"), opts), 2, 0, lib=lib)
        chunk = 50
        for k in range(steps // chunk):
            r.tape(clear=True)
            r.simulate(chunk)
            tape = r.tape(clear=True)
            for rep in (0, 1):
                sim.engine.attach_tape(rep, tape)
            sim.engine.run(chunk, 0, 0, 0)
            sim.engine.assert_ok()
            for rep in (0, 1):
                assert sim.engine.tape_position(rep) == len(tape), (system, moveset, k)
                assert_state_equal(sim.engine.state(rep), r.state(), f"{system} {moveset} chunk {k}")
            got = sim.engine.order_params()[0]
            for i, tag in enumerate(TAGS):
                # stored values: evaluating a per-domain parameter in the oracle would refresh it and change its future
                want, defined = r.order_param_stored(tag)
                if tag in ("numstaples", "numfulldomains", "dist-mid"):
                    want = r.order_param(tag)  # move-update kind: the engine's getter re-evaluates these, as the reference's does
                assert got[i] == want, (system, moveset, k, tag, list(got))
                undefined_seen |= not defined
            assert abs(sim.engine.energies()[0, 4] - r.total_bias_stored()) < 1e-12, (system, moveset, k)
            seen.add((int(got[3]), int(got[4]), round(r.total_bias_stored(), 6)))
        att, acc = sim.engine.move_stats()
        ra, rb = r.move_stats()
        assert list(att[0]) == list(ra) and list(acc[0]) == list(rb)
    assert len(seen) > 10  # the per-domain parameters and biases actually move
    assert undefined_seen  # ... and the stale "undefined" state the reference leaves behind was part of the trajectories


def test_per_domain_biases_replay(hostsim_lib, oracle, tmp_path):
    per_domain_biases(hostsim_lib, oracle, tmp_path)


def test_per_domain_biases_with_whole_system_passes(hostsim_lib, oracle, tmp_path):
    """centering and check_all_constraints go through the same virtual set / unassign calls (origami_system.cpp:267-325,
    553-571): the bias bookkeeping they leave behind (biases zeroed by set_domain_config's stale parameters) is part of
    the trajectory."""
    opts_extra = dict(centering_freq=40, constraint_check_freq=70)
    cases = [("snodin_unbound.json", "moveset_standard.json", 339, 21, 1000)]
    seen = set()
    for system, moveset, temp, seed, steps in cases:
        opts = make_options(system, moveset, temp=temp, domain_update_biases_present=True,
                            bias_functions_file=os.path.join(INPUTS, "biases_perdomain.json"), **opts_extra)
        opts["order_parameter_file"] = os.path.join(INPUTS, "ops_perdomain.json")
        r = oracle.RefSystem(opts)
        r.seed(seed)
        r.simulate(steps)
        sim = Simulation(write_inp(str(tmp_path / "pdw.inp"), opts), 1, 0, lib=hostsim_lib)
        sim.engine.attach_tape(0, r.tape())
        sim.engine.run(steps, 40, 0, 70)
        sim.engine.assert_ok()
        assert_state_equal(sim.engine.state(0), r.state())
        assert abs(sim.engine.energies()[0, 4] - r.total_bias_stored()) < 1e-12


@pytest.mark.gpu
def test_per_domain_biases_replay_gpu(oracle, tmp_path):
    per_domain_biases(None, oracle, tmp_path)
