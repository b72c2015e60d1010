"""Output / restart surface on an EVOLVING trajectory.

* The host driver fed with the tape of a live run of the unmodified reference (outputs enabled) must write the same
  .trj .vcf .states .ores .counts .staples .staplestates .ene .ops files byte for byte, and the same .moves
  counters (files.cpp:519-793, simulation.cpp:641-646, 706-718).
* restart_traj_file / restart_step (files.cpp:129-218, origami_system.cpp:1000-1003): a trajectory written by the
  reference is read back by both codes, which then continue identically under a replayed tape.
* Replica-exchange restart (ptmc_simulation.cpp:38-83): per-replica restart_traj_filebase-<rank><postfix> and the last
  row of restart_swap_file.
* rand_engine_state_output_freq / read_rand_engine_state (files.cpp:220-246, 781-793; simulation.cpp:204-212): the state
  of the generator written next to the trajectory; a run restarted from frame k of both continues as the uninterrupted
  run did (production draws, Philox state instead of the reference's mt19937_64 text).
CPU: host emulation of the device sources; GPU: the CUDA library."""
import numpy as np
import pytest

from conftest import assert_state_equal, make_options, write_inp
from latticednaorigami_b200.binding import Simulation

EXTS = [".trj", ".vsf", ".vcf", ".states", ".ores", ".counts", ".staples", ".staplestates", ".ene", ".ops"]
OUT = dict(configs_output_freq=100, vtf_output_freq=100, counts_output_freq=100, order_params_output_freq=100, energies_output_freq=100,
           ops_to_output="numstaples numfulldomains nummisdomains numstackedpairs")


def evolving_outputs(lib, oracle, tmp_path, system="snodin_unbound.json", temp=337, steps=800, seed=21):
    ref_opts = make_options(system, temp=temp, ct_steps=steps, output_filebase=str(tmp_path / "ref"), **OUT)
    r = oracle.RefSystem(ref_opts, workdir=str(tmp_path))
    r.seed(seed)
    r.simulate(steps)
    tape = r.tape()
    final = r.state()
    att, acc = r.move_stats()
    r.close()  # closes the reference's output files
    our_opts = make_options(system, temp=temp, ct_steps=steps, random_seed=1, output_filebase=str(tmp_path / "our"), **OUT)
    sim = Simulation(write_inp(str(tmp_path / "our.inp"), our_opts), 1, 0, lib=lib)
    sim.engine.attach_tape(0, tape)
    sim.run()
    sim.engine.assert_ok()
    assert sim.engine.tape_position(0) == len(tape)
    assert_state_equal(sim.engine.state(0), final)
    for ext in EXTS:
        ref, our = (tmp_path / ("ref" + ext)).read_text(), (tmp_path / ("our" + ext)).read_text()
        assert ref, ext
        if ext == ".ene":
            # energies are running fp64 sums printed with 10 digits: equal to 1e-12 of their terms, not bitwise (the
            # stacking column of an unstacked system is the rounding residue energy - (enthalpy - entropy) itself)
            assert our.splitlines()[0] == ref.splitlines()[0]
            a, b = np.loadtxt(tmp_path / "ref.ene", skiprows=1), np.loadtxt(tmp_path / "our.ene", skiprows=1)
            assert a.shape == b.shape and np.array_equal(a[:, 0], b[:, 0])
            assert np.all(np.abs(a - b) <= 1e-9 * np.maximum(1.0, np.abs(a).max(axis=1, keepdims=True)))
            continue
        assert our == ref, f"{ext} differs from the reference's file"
    # the trajectory evolves: frames differ and staples come and go
    frames = [f for f in (tmp_path / "our.trj").read_text().split("\n\n") if f.strip()]
    assert len(frames) == steps // 100 and len({f.split("\n", 1)[1] for f in frames}) == len(frames)
    counts = np.loadtxt(tmp_path / "our.counts")
    assert len(set(counts[:, 1])) > 1
    moves = (tmp_path / "our.moves").read_text()
    for i, label in enumerate(sim.movetype_labels):
        if label != "Orientation rotation":  # (the reference's summary leaves it out, orientation_movetype.cpp:28)
            assert f"Movetype: {label}\n    Attempts: {att[i]}\n    Accepts: {acc[i]}\n" in moves


def moves_summary(lib, oracle, tmp_path, system, moveset, temp, steps, seed):
    """<filebase>.moves (simulation.cpp:706-718) with every movetype's own breakdown from its typed trackers - staple
    moves by staple type, scaffold regrowth by segment length (and staples) - byte for byte against the file the reference
    CLI writes for the same run; the draws of that run are taped from the oracle library seeded alike."""
    import subprocess
    kw = dict(temp=temp, ct_steps=steps, configs_output_freq=steps // 2)
    ref_inp = write_inp(str(tmp_path / "ref.inp"), make_options(system, moveset, random_seed=seed, output_filebase=str(tmp_path / "ref"), **kw))
    subprocess.run([oracle.CLI_PATH, "-i", ref_inp], check=True, capture_output=True)
    r = oracle.RefSystem(make_options(system, moveset, **kw))
    r.seed(seed)
    r.simulate(steps)
    tape = r.tape()
    sim = Simulation(write_inp(str(tmp_path / "our.inp"), make_options(system, moveset, random_seed=1, output_filebase=str(tmp_path / "our"), **kw)), 1, 0, lib=lib)
    sim.engine.attach_tape(0, tape)
    sim.run()
    sim.engine.assert_ok()
    assert (tmp_path / "our.trj").read_text() == (tmp_path / "ref.trj").read_text()  # the same run
    ours, ref = (tmp_path / "our.moves").read_text(), (tmp_path / "ref.moves").read_text()
    assert ours == ref
    return ref


def test_moves_summary_matches_reference_cli(hostsim_lib, oracle, tmp_path):
    for k, (system, moveset, temp, steps, seed, needles) in enumerate([
            ("snodin_assembled.json", "moveset_standard.json", 338, 3000, 5, ["Insertion attempts", "Number of scaffold domains: 12", "Staple type: 12"]),
            ("snodin_unbound.json", "moveset_ctcb.json", 334, 2500, 6, ["Number of staples: 0", "Number of scaffold domains"]),
            ("four_unbound.json", "moveset_four.json", 350, 2000, 7, ["Staple type"]),
            # the transform / linker movetypes: three tables keyed by pairs (transform_movetypes.cpp:64-139)
            ("snodin_assembled.json", "moveset_linker.json", 340, 2500, 8, ["Number of linker/central domains", "Number of linker/central staples",
                                                                            "Sum/number of displacement/turns"]),
            ("snodin_unbound.json", "moveset_linker_heavy.json", 336, 2000, 9, ["Sum/number of displacement/turns: -"])]):
        d = tmp_path / str(k)
        d.mkdir()
        text = moves_summary(hostsim_lib, oracle, d, system, moveset, temp, steps, seed)
        for needle in needles:
            assert needle in text, (system, needle)


@pytest.mark.gpu
def test_moves_summary_matches_reference_cli_gpu(oracle, tmp_path):
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    moves_summary(None, oracle, tmp_path / "a", "snodin_assembled.json", "moveset_standard.json", 338, 3000, 5)
    text = moves_summary(None, oracle, tmp_path / "b", "snodin_assembled.json", "moveset_linker.json", 340, 2500, 8)
    assert "Sum/number of displacement/turns" in text


def trj_restart(lib, oracle, tmp_path):
    # a trajectory written by the reference ...
    a_opts = make_options("snodin_unbound.json", temp=336, output_filebase=str(tmp_path / "a"), configs_output_freq=150)
    a = oracle.RefSystem(a_opts, workdir=str(tmp_path))
    a.seed(5)
    a.simulate(600)
    frame3 = a.state()  # the 4th frame (restart_step counts frames from 0)
    a.close()
    # ... is read back by the reference and by the host driver; both continue under the same tape
    kw = dict(temp=338, restart_traj_file=str(tmp_path / "a.trj"), restart_step=3)
    b = oracle.RefSystem(make_options("snodin_unbound.json", output_filebase=str(tmp_path / "b"), **kw), workdir=str(tmp_path))
    assert_state_equal(b.state(), frame3, "reference restart")
    sim = Simulation(write_inp(str(tmp_path / "r.inp"), make_options("snodin_unbound.json", random_seed=1, **kw)), 2, 0, lib=lib)
    for rep in (0, 1):
        assert_state_equal(sim.engine.state(rep), frame3, "restart")
    e = b.energy()
    assert abs(sim.engine.energies()[0, 0] - e) <= 1e-12 * max(1.0, abs(e))
    b.seed(6)
    b.simulate(300)
    tape = b.tape()
    for rep in (0, 1):
        sim.engine.attach_tape(rep, tape)
    sim.engine.run(300)
    sim.engine.assert_ok()
    for rep in (0, 1):
        assert_state_equal(sim.engine.state(rep), b.state(), "continued")
    # a missing frame is an error in both (files.cpp:203-206)
    with pytest.raises(Exception):
        Simulation(write_inp(str(tmp_path / "bad.inp"), make_options("snodin_unbound.json", restart_traj_file=str(tmp_path / "a.trj"),
                                                                     restart_step=9)), 1, 0, lib=lib)


def randstate_restart(lib, tmp_path):
    """An interrupted run = the uninterrupted one: restart from frame 2 of .trj and line 2 of .randstate (production mode).
    The key, the subsequence, the draw counter and the buffered words all come from the file: no seed is given."""
    kw = dict(temp=338, configs_output_freq=100, rand_engine_state_output_freq=100)
    a = Simulation(write_inp(str(tmp_path / "a.inp"), make_options("snodin_unbound.json", random_seed=77, ct_steps=600,
                                                                    output_filebase=str(tmp_path / "a"), **kw)), 1, 0, lib=lib)
    a.run()
    a.engine.assert_ok()
    lines = (tmp_path / "a.randstate").read_text().splitlines()
    W = a.engine.rng_state().shape[1]
    assert len(lines) == 6 and all(len(l.split()) == W for l in lines)
    assert [int(x) for x in lines[-1].split()] == a.engine.rng_state()[0].tolist()
    assert int(lines[0].split()[0]) == 77 and len({l.split()[4] for l in lines}) == 6  # key = seed; the counter advances
    b = Simulation(write_inp(str(tmp_path / "b.inp"), make_options(
        "snodin_unbound.json", ct_steps=300, restart_traj_file=str(tmp_path / "a.trj"), restart_step=2, read_rand_engine_state=True,
        rand_engine_state_file=str(tmp_path / "a.randstate"), output_filebase=str(tmp_path / "b"), **kw)), 1, 0, lib=lib)
    assert [int(x) for x in lines[2].split()] == b.engine.rng_state()[0].tolist()
    b.run()
    b.engine.assert_ok()
    # (unique chain indices restart at the largest one of the frame, origami_system.cpp:1000-1003: everything else is equal)
    def without_uids(st):  # bound partners name their chain by its unique index: renamed to the chain's position
        st = dict(st)
        rank = {int(u): k for k, u in enumerate(st["chain_index"])}
        bound = st["bound"].copy()
        for row in bound:
            if row[0] >= 0:
                row[0] = rank[int(row[0])]
        st["bound"] = bound
        st["chain_index"] = np.arange(len(st["chain_index"]), dtype=st["chain_index"].dtype)
        return st
    assert_state_equal(without_uids(b.engine.state(0)), without_uids(a.engine.state(0)), "restarted run")
    assert np.array_equal(b.engine.rng_state(), a.engine.rng_state())

    def frames(name):  # chain lines without the unique index
        out = []
        for f in (tmp_path / name).read_text().split("\n\n"):
            if f.strip():
                rows = f.strip().split("\n")[1:]
                out.append(tuple(r.split()[1] if k % 3 == 0 else r for k, r in enumerate(rows)))
        return out
    fa, fb = frames("a.trj"), frames("b.trj")
    assert fa[3:] == fb and len(set(fa)) == 6  # the same frames after the restart point, on an evolving trajectory
    assert (tmp_path / "b.randstate").read_text().splitlines() == lines[3:]
    # a specified seed wins over the file (simulation.cpp:200-204); a missing line or file is a FileError
    c = Simulation(write_inp(str(tmp_path / "c.inp"), make_options(
        "snodin_unbound.json", random_seed=5, read_rand_engine_state=True, rand_engine_state_file=str(tmp_path / "a.randstate"), **kw)), 1, 0, lib=lib)
    assert c.engine.rng_state()[0, 0] == 5 and c.engine.rng_state()[0, 4] == 0
    for bad in (dict(restart_step=9, rand_engine_state_file=str(tmp_path / "a.randstate")), dict(rand_engine_state_file=str(tmp_path / "none.randstate"))):
        with pytest.raises(Exception, match="not found|does not exist"):
            Simulation(write_inp(str(tmp_path / "bad.inp"), make_options("snodin_unbound.json", read_rand_engine_state=True, **bad)), 1, 0, lib=lib)
    # a line that is not a Philox state of this engine (here: the text form of another generator) is refused
    (tmp_path / "mt.randstate").write_text(" ".join(str(2 ** 40 + k) for k in range(313)) + "\n")
    with pytest.raises(Exception, match="not a Philox state"):
        Simulation(write_inp(str(tmp_path / "bad.inp"), make_options("snodin_unbound.json", read_rand_engine_state=True,
                                                                      rand_engine_state_file=str(tmp_path / "mt.randstate"))), 1, 0, lib=lib)
    # a batch of replicas writes and reads <filebase>-<replica>.randstate
    d = Simulation(write_inp(str(tmp_path / "d.inp"), make_options("snodin_unbound.json", random_seed=9, ct_steps=200,
                                                                    output_filebase=str(tmp_path / "d"), **kw)), 2, 0, lib=lib)
    d.run()
    e = Simulation(write_inp(str(tmp_path / "e.inp"), make_options(
        "snodin_unbound.json", restart_step=1, read_rand_engine_state=True, rand_engine_state_file=str(tmp_path / "d.randstate"))), 2, 0, lib=lib)
    assert np.array_equal(e.engine.rng_state(), d.engine.rng_state()) and e.engine.rng_state()[1, 2] == 1


def test_randstate_restart(hostsim_lib, tmp_path):
    randstate_restart(hostsim_lib, tmp_path)


def test_output_files_on_an_evolving_trajectory(hostsim_lib, oracle, tmp_path):
    evolving_outputs(hostsim_lib, oracle, tmp_path)


def test_trj_restart(hostsim_lib, oracle, tmp_path):
    trj_restart(hostsim_lib, oracle, tmp_path)


def test_replica_exchange_restart(hostsim_lib, oracle, tmp_path):
    from test_exchange_oracle import exchange_against_oracle
    exchange_against_oracle(oracle, tmp_path, hostsim_lib, "ut", swaps=8, restart=True)


@pytest.mark.gpu
def test_outputs_and_restart_gpu(oracle, tmp_path):
    from test_exchange_oracle import exchange_against_oracle
    (tmp_path / "o").mkdir()
    (tmp_path / "r").mkdir()
    (tmp_path / "p").mkdir()
    evolving_outputs(None, oracle, tmp_path / "o")
    trj_restart(None, oracle, tmp_path / "r")
    (tmp_path / "s").mkdir()
    randstate_restart(None, tmp_path / "s")
    exchange_against_oracle(oracle, tmp_path / "p", None, "ut", swaps=8, restart=True)


def _log_blocks(text):
    out, cur = [], None
    for line in text.splitlines():
        if line.startswith("Step: "):
            cur = [line]
        elif cur is not None:
            if line == "":
                out.append("\n".join(cur))
                cur = None
            else:
                cur.append(line)
    return out


def test_log_entries_match_reference_cli(hostsim_lib, oracle, tmp_path, capfd):
    """logging_freq (simulation.cpp:628-630, 667-705): the entries a constant-temperature run prints to stdout - counters,
    staple counts, energy, biases, the movetype of the reported step and whether it was accepted - byte for byte against the
    reference CLI on the same (replayed) run."""
    import subprocess
    kw = dict(temp=338, ct_steps=700, logging_freq=100, configs_output_freq=350)
    ref_inp = write_inp(str(tmp_path / "ref.inp"), make_options("snodin_assembled.json", random_seed=5, output_filebase=str(tmp_path / "ref"), **kw))
    ref_out = subprocess.run([oracle.CLI_PATH, "-i", ref_inp], check=True, capture_output=True, text=True).stdout
    r = oracle.RefSystem(make_options("snodin_assembled.json", **kw))
    r.seed(5)
    r.simulate(700)
    sim = Simulation(write_inp(str(tmp_path / "our.inp"), make_options("snodin_assembled.json", random_seed=1, output_filebase=str(tmp_path / "our"), **kw)),
                     1, 0, lib=hostsim_lib)
    sim.engine.attach_tape(0, r.tape())
    capfd.readouterr()
    sim.run()
    ours = capfd.readouterr().out
    assert (tmp_path / "our.trj").read_text() == (tmp_path / "ref.trj").read_text()  # the same run
    a, b = _log_blocks(ref_out), _log_blocks(ours)
    assert len(a) == 7 and a == b
    assert len({blk.split("Movetype: ")[1] for blk in a}) > 2  # several movetypes and both outcomes are reported


def test_replica_exchange_logs_to_out_files(hostsim_lib, tmp_path):
    """The replica-exchange drivers log to <filebase>-<rank>.out (ptmc_simulation.cpp:56): one entry per logging_freq
    steps and replica, with the temperature the replica holds at that step."""
    opts = make_options("snodin_unbound.json", simulation_type="ut_parallel_tempering", random_seed=3, temps=[330, 335, 340], num_reps=3,
                        chem_pot_mults=[1, 1, 1], bias_mults=[1, 1, 1], stacking_mults=[1, 1, 1], exchange_interval=50, swaps=6,
                        max_pt_dur=1e9, logging_freq=50, output_filebase=str(tmp_path / "pt"))
    sim = Simulation(write_inp(str(tmp_path / "pt.inp"), opts), 3, 0, lib=hostsim_lib)
    sim.run()
    temps_seen = []
    for k in range(3):
        blocks = _log_blocks((tmp_path / f"pt-{k}.out").read_text())
        assert [b.splitlines()[0] for b in blocks] == [f"Step: {50 * (i + 1)}" for i in range(6)]
        assert all(len(b.splitlines()) == 16 for b in blocks)
        temps_seen.append([b.splitlines()[1] for b in blocks])
    # at every logged step the three replicas hold the three temperatures of the ladder
    for i in range(6):
        assert sorted(t[i] for t in temps_seen) == ["Temperature: 330", "Temperature: 335", "Temperature: 340"]
