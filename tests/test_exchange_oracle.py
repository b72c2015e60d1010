"""Replica exchange against the reference's OWN drivers (row a26): the unmodified PTGCMCSimulation subclasses
(ptmc_simulation.cpp:106-150, 360-412, 495-560, 651-680) run in the oracle with one thread per rank; every rank's
draws are taped, rank 0's tape also carries the master's exchange draws (App. A20). The engine replays the MC
tapes on three replicas and the exchange draws through ldo_set_exchange_tape, and must reproduce

  * the .swp permutation sequence row by row (ptmc_simulation.cpp:315-322),
  * every replica's final lattice state bit for bit and its energy to 1e-12,
  * every replica's move statistics,

for ut_ (examples/ptmc.inp), t_, st_, hut_ and 2d_parallel_tempering. CPU: host emulation; GPU: the CUDA library."""
import os

import numpy as np
import pytest

from conftest import INPUTS, assert_state_equal, make_options, write_inp
from latticednaorigami_b200.binding import Simulation

CASES = {
    "ut": dict(simulation_type="ut_parallel_tempering", temps=[330.0, 333.0, 336.0], chem_pot_mults=[1, 1, 1], bias_mults=[1, 1, 1],
               stacking_mults=[1, 1, 1]),
    "t": dict(simulation_type="t_parallel_tempering", temps=[331.0, 334.0, 338.0], chem_pot_mults=[1, 1, 1], bias_mults=[1, 1, 1],
              stacking_mults=[1, 1, 1]),
    "st": dict(simulation_type="st_parallel_tempering", temps=[330.0, 332.0, 334.0], chem_pot_mults=[1, 1, 1], bias_mults=[1, 1, 1],
               stacking_mults=[1.0, 0.9, 0.8]),
    "hut": dict(simulation_type="hut_parallel_tempering", temps=[332.0, 335.0, 338.0], chem_pot_mults=[1.0, 1.1, 1.2],
                bias_mults=[1.0, 0.6, 0.2], stacking_mults=[1, 1, 1], bias_functions_file=os.path.join(INPUTS, "biases_dist.json"),
                order_parameter_file=os.path.join(INPUTS, "ops_dist.json")),
    "2d": dict(simulation_type="2d_parallel_tempering", temps=[331.0, 336.0], stacking_mults=[1.0, 0.85], num_reps=4),
}


def exchange_against_oracle(oracle, tmp_path, lib, case, system="snodin_unbound.json", swaps=12, interval=150, seed0=900, restart=False):
    extra = dict(CASES[case])
    n = extra.pop("num_reps", 3)
    if restart:
        # a first run of the reference writes per-replica trajectories and the swap file; the compared run restarts from
        # the configurations of its 4th frame and from the last row of its .swp (ptmc_simulation.cpp:38-83)
        first = make_options(system, num_reps=n, swaps=6, exchange_interval=interval, max_pt_dur=1e9, configs_output_freq=interval,
                             restart_from_swap=False, output_filebase=str(tmp_path / "first"), **extra)
        oracle.pt_run(first, n, [seed0 + 5 + r for r in range(n)], record_tapes=False, workdir=str(tmp_path))
        extra.update(restart_from_config=True, restart_traj_filebase=str(tmp_path / "first"), restart_traj_postfix=".trj", restart_step=3,
                     restart_swap_file=str(tmp_path / "first.swp"))
    opts = make_options(system, num_reps=n, swaps=swaps, exchange_interval=interval, max_pt_dur=1e9, configs_output_freq=interval,
                        restart_from_swap=restart, **extra)
    ref_opts = dict(opts, output_filebase=str(tmp_path / f"ref_{case}"))
    ref = oracle.pt_run(ref_opts, n, [seed0 + 17 * r for r in range(n)], workdir=str(tmp_path))
    assert len(ref["swp"]) == swaps + 1 and (restart or ref["swp"][0] == list(range(n)))
    assert len(ref["marks"]) == swaps

    ours = dict(opts, random_seed=1, output_filebase="")
    sim = Simulation(write_inp(str(tmp_path / f"our_{case}.inp"), ours), n, 0, lib=lib)
    eng = sim.engine
    for r in range(n):
        eng.attach_tape(r, ref["mc_tapes"][r])
    eng.set_exchange_tape(ref["exchange_reals"], ref["exchange_offsets"])
    two_d = case == "2d"
    shifted_rounds = 0
    for swap_i in range(1, swaps + 1):
        assert sim.exchange_advance() == 0
        if swap_i == swaps:
            # the reference's last round ends without another update_control_qs(): its final energies are those
            # of the control variables the round was run with
            e = eng.energies()[:, 0]
            for r in range(n):
                assert abs(e[r] - ref["energies"][r]) <= 1e-12 * max(1.0, abs(ref["energies"][r])), (case, r)
        # A swap test whose probability is 1 up to rounding consumes a draw in one code and not in the other: the
        # reference's running energies carry the rounding residue of its weight passes, the device restores the energy from
        # a snapshot there, so p = exp(1e-13) on one side and exp(-1e-13) on the other. The test itself comes
        # out the same (accepted unless u > 1 - 1e-13), but the next pair of that round reads a shifted tape. Such a round
        # shows in the tape status (draws missing / unused); only such a round may decide differently, and the comparison
        # ends there (the ladders have parted).
        before = sum(eng.exchange_tape_status())
        sim.exchange_apply(swap_i)
        shifted = sum(eng.exchange_tape_status()) != before
        shifted_rounds += shifted
        q2r = sim.exchange_state(1, n, two_d=two_d)[0][0]
        if list(q2r) != ref["swp"][swap_i]:
            assert shifted, (case, swap_i, list(q2r), ref["swp"][swap_i])
            assert swap_i > 1 and sorted(q2r) == list(range(n))
            print(f"{case}: a test of round {swap_i} was decided by rounding; compared up to round {swap_i - 1}")
            return len(ref["exchange_reals"]), len({tuple(p) for p in ref["swp"]})
    eng.assert_ok()
    # draws one code made and the other did not (a probability rounding to exactly 1 in one of them only; the
    # running energies agree to 1e-12, not bitwise; replicas in the same binding state at different temperatures meet
    # this at every round): no decision compared above depended on them
    missing, unused = eng.exchange_tape_status()
    assert missing + unused <= n * shifted_rounds, (missing, unused, shifted_rounds)
    att, acc = eng.move_stats()
    for r in range(n):
        assert eng.tape_position(r) == len(ref["mc_tapes"][r]), (case, r)
        got = eng.state(r)
        for k in ("chain_index", "chain_ident", "chain_len", "pos", "ore"):
            assert np.array_equal(got[k], ref["states"][r][k]), (case, r, k)
        assert list(att[r]) == list(ref["attempts"][r]) and list(acc[r]) == list(ref["accepts"][r])
    # the test must have exercised both outcomes with a real draw
    perms = {tuple(p) for p in ref["swp"]}
    return len(ref["exchange_reals"]), len(perms)


@pytest.mark.parametrize("case", list(CASES))
def test_swap_sequence_matches_reference_driver(hostsim_lib, oracle, tmp_path, case):
    draws, perms = exchange_against_oracle(oracle, tmp_path, hostsim_lib, case)
    assert perms >= 2


def test_exchange_draws_are_consumed(hostsim_lib, oracle, tmp_path):
    # over the five variants at least some swap tests are decided by a draw (p < 1), both ways
    total = sum(exchange_against_oracle(oracle, tmp_path / c, hostsim_lib, c, seed0=4000)[0] for c in CASES if (tmp_path / c).mkdir() is None)
    assert total >= 5


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_swap_sequence_matches_reference_driver_gpu(oracle, tmp_path, case):
    exchange_against_oracle(oracle, tmp_path, None, case)
