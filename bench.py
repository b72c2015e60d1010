#!/usr/bin/env python
"""bench.py — attempted MC moves / second (whole job) of the replica-batched engine, next to the reference's own
CPU code on the host cores.

Default workload `ptmc` (BASELINE.json configs[2]): examples/ptmc.inp generalised to a batch:
`ut_parallel_tempering`, exchange_interval 100, 32-temperature ladder 330..361 K (1 K steps), 16384 replicas per GPU
= 512*N ladders, ladder slots dealt in serpentine order over the N GPUs, moveset_standard. The ensemble is
STATIONARY: every ladder starts from the same equilibrated 32-slot ladder (bench_data/snodin_ladder32.json: 200 000
moves per replica of the unmodified reference's own replica-exchange driver, bench_data/make_ladder.py), tiled over
the ladders with distinct Philox subsequences; the warm-up rounds decorrelate the copies. One "step" = one exchange
round: 100 attempted moves on every replica, collection of the exchange records, (N > 1: NCCL all-gather), on-device
swap decisions and the energy rebuild that follows a control-variable update. The reference arm restarts the
unmodified reference CLI from the SAME ladder states (.trj restart files), one process per host core at the slot's
temperature.

Other workloads (`--workload`): ct_four (configs[1]), ptmwus (configs[3]), anneal_large (configs[4]); see WORKLOADS.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code, host cores
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

os.environ.setdefault("LDO_QUIET", "1")  # the host library must not write to stdout: the bench prints ONE JSON line

# Native libraries (NCCL's version banner, the C++ host) write to file descriptor 1 directly: keep the real
# stdout for the single JSON line and point fd 1 at stderr for everything else.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(obj) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
INPUTS = os.path.join(ROOT, "tests", "golden", "inputs")
BENCH_DATA = os.path.join(ROOT, "bench_data")

LADDER = [330.0 + i for i in range(32)]
EXCHANGE_INTERVAL = 100
SMEM_BYTES_PER_MOVE = 41.7e3  # 2215 table ops * 8 B + 1500 record touches * 16 B
HBM_FALLBACK_GBS = 6650.0  # B200_PROFILING.md fallback
REF_MOVES_PER_PROC = 20000  # bounded sample of the reference arm: moves per process per step


def base_options():
    return {
        "origami_input_filename": os.path.join(INPUTS, "snodin_unbound.json"),
        "domain_type": "HalfTurn", "binding_pot": "FourBody", "misbinding_pot": "Opposing",
        "stacking_pot": "Constant", "hybridization_pot": "NearestNeighbour", "apply_mean_field_cor": "false",
        "staple_M": 1e-7, "cation_M": 0.5, "staple_u_mult": 1, "stacking_ene": -1000,
        "max_total_staples": 24, "max_type_staples": 12, "max_staple_size": 2,
        "domain_update_biases_present": "false",
        "order_parameter_file": os.path.join(INPUTS, "ops_standard.json"),
        "movetype_file": os.path.join(INPUTS, "moveset_standard.json"),
        "centering_freq": 100000, "constraint_check_freq": 1000000,
    }


def write_inp(path, options):
    with open(path, "w") as f:
        for k, v in options.items():
            if isinstance(v, (list, tuple)):
                v = " ".join(str(x) for x in v)
            f.write(f"{k}={v}\n")
    return path


def write_trj(path, state, step=0):
    """One frame in the reference's trajectory format (OrigamiTrajOutputFile::write, files.cpp:529-548), readable by
    OrigamiTrajInputFile::read_config (files.cpp:129-218) as restart_traj_file / restart_step 0."""
    with open(path, "w") as f:
        f.write(f"{step}\n")
        k = 0
        for ci, cid, cl in zip(state["chain_index"], state["chain_ident"], state["chain_len"]):
            f.write(f"{ci} {cid}\n")
            f.write("".join(f"{v} " for v in state["pos"][3 * k:3 * (k + cl)]) + "\n")
            f.write("".join(f"{v} " for v in state["ore"][3 * k:3 * (k + cl)]) + "\n")
            k += cl
        f.write("\n")
    return path


def load_states(name):
    path = os.path.join(BENCH_DATA, name)
    if not os.path.exists(path):
        raise SystemExit(f"bench.py: {path} missing (see bench_data/README.md for the generating script)")
    return json.load(open(path))


# ---------------------------------------------------------------------------------------------
# Workloads
# ---------------------------------------------------------------------------------------------

class Workload:
    """What both arms share: the option set, the committed starting states and how a replica maps onto them."""
    name = ""
    metric = "attempted MC moves/sec (whole box)"
    default_replicas = 16384
    moves_per_step = EXCHANGE_INTERVAL  # per replica

    def config(self, n_gpus, R):
        raise NotImplementedError


class PTMC(Workload):
    name = "ptmc"
    metric = "attempted MC moves/sec (whole box), snodin PTMC"

    def __init__(self):
        self.data = load_states("snodin_ladder32.json")
        assert self.data["ladder"] == LADDER

    def config(self, n_gpus, R):
        return {"workload": "snodin ut_parallel_tempering (examples/ptmc.inp batched): 32-temperature ladder 330..361 K, "
                            "exchange_interval 100, moveset_standard; stationary start: bench_data/snodin_ladder32.json "
                            "(equilibrated by the reference's own PT driver, 200000 moves/replica), tiled over the ladders",
                "replicas_per_gpu": R, "ladders": R // len(LADDER) * n_gpus, "ladder_len": len(LADDER),
                "moves_per_step": R * n_gpus * EXCHANGE_INTERVAL,
                "l2_policy": f"state ({R * 3.3e3 / 1e6:.0f} MB/GPU) is re-staged from HBM every launch; per-step working set is shared memory"}

    def options(self):
        L = len(LADDER)
        o = base_options()
        o.update({"simulation_type": "ut_parallel_tempering", "num_reps": L, "temps": LADDER, "chem_pot_mults": [1] * L,
                  "bias_mults": [1] * L, "stacking_mults": [1] * L, "exchange_interval": EXCHANGE_INTERVAL, "swaps": 0,
                  "random_seed": 20261017})
        return o

    def reference_job(self, proc, step_k, n_procs):
        """(options, starting state) of reference process `proc` in step `step_k`: the ladder slots are covered in turn."""
        slot = (proc + step_k * n_procs) % len(LADDER)
        o = base_options()
        o.update({"simulation_type": "constant_temp", "temp": LADDER[slot]})
        return o, self.data["slots"][slot]


class StatesWorkload(Workload):
    """Independent replicas of one parameter set, started from committed states generated by the reference itself
    (bench_data/make_states.py); replica r starts from state r % n_states."""
    states_file = ""

    def __init__(self):
        self.data = load_states(self.states_file)

    def reference_job(self, proc, step_k, n_procs):
        i = (proc + step_k * n_procs) % len(self.data["states"])
        st = self.data["states"][i]
        return self.state_options(st), st

    def state_options(self, st):
        return self.options()


class CTFour(StatesWorkload):
    name = "ct_four"
    metric = "attempted MC moves/sec (whole box), four_unbound constant-T"
    states_file = "ct_four_states.json"
    moves_per_step = 1000

    def config(self, n_gpus, R):
        return {"workload": "four_unbound constant-temperature replicas at 330 K (examples/enum.inp system and limits, "
                            "moveset_four), independent Philox streams; stationary start: bench_data/ct_four_states.json",
                "replicas_per_gpu": R, "moves_per_step": R * n_gpus * self.moves_per_step,
                "l2_policy": "state is re-staged from HBM every launch; per-step working set is shared memory"}

    def options(self):
        o = base_options()
        o.update({"origami_input_filename": os.path.join(INPUTS, "four_unbound.json"), "movetype_file": os.path.join(INPUTS, "moveset_four.json"),
                  "max_total_staples": 2, "max_type_staples": 2, "simulation_type": "constant_temp", "temp": 330, "random_seed": 20261018})
        return o


WINDOWS = [(2 * k, 2 * k + 4) for k in range(11)]  # width-4 stride-2 windows over numfulldomains 0..24


class PTMWUS(StatesWorkload):
    name = "ptmwus"
    metric = "attempted MC moves/sec (whole box), snodin PTMWUS"
    states_file = "ptmwus_states.json"
    default_replicas = 11 * 1488

    def config(self, n_gpus, R):
        return {"workload": "snodin ptmw_umbrella_sampling (examples/ptmwus.inp batched): 11 windows of width 4, stride 2 on "
                            "numfulldomains at 354 K, exchange_interval 100, window exchange every step; stationary start: "
                            "bench_data/ptmwus_states.json",
                "replicas_per_gpu": R, "windows": len(WINDOWS), "ladders": R // len(WINDOWS) * n_gpus,
                "moves_per_step": R * n_gpus * EXCHANGE_INTERVAL,
                "l2_policy": "state is re-staged from HBM every launch; per-step working set is shared memory"}

    def options(self, tmp=None):
        o = base_options()
        o.update({"temp": 354, "bias_functions_file": os.path.join(INPUTS, "biases_mwus-numfulldomains.json"), "bias_functions_mult": 1,
                  "simulation_type": "ptmw_umbrella_sampling", "us_grid_bias_tag": "grid", "max_num_iters": 1, "max_D_bias": 10,
                  "exchange_interval": EXCHANGE_INTERVAL, "equil_steps": 0, "iter_steps": 0, "iter_swaps": 0, "multi_window": "true",
                  "random_seed": 20261019})
        if tmp is not None:
            wf = os.path.join(tmp, "bench.windows")
            with open(wf, "w") as f:
                f.write("lswnumfulldomains\n" + "".join(f"{a}, {b}\n" for a, b in WINDOWS))
            o["windows_file"] = wf
        return o

    def state_options(self, st):
        # one reference process = one window: constant-T with the window's restraint written into its own bias file
        o = base_options()
        o.update({"temp": 354, "simulation_type": "constant_temp", "bias_functions_mult": 1, "_window": st["window"]})
        return o


class AnnealLarge(StatesWorkload):
    name = "anneal_large"
    metric = "attempted MC moves/sec (whole box), 168-domain ThreeQuarterTurn annealing"
    states_file = "anneal_large_states.json"
    default_replicas = 4096
    moves_per_step = 100

    def config(self, n_gpus, R):
        d = self.data
        return {"workload": f"synthetic 168-domain ThreeQuarterTurn raster (12 x 14, 84 staple types, Uniform potential), annealing "
                            f"sweep {d['max_temp']:.0f} -> {d['min_temp']:.0f} K; the timed steps are constant-T stages at the cold end "
                            f"({d['bench_temp']:.0f} K) from bench_data/anneal_large_states.json (end of the reference's own sweep)",
                "replicas_per_gpu": R, "moves_per_step": R * n_gpus * self.moves_per_step,
                "l2_policy": f"replica state ({R * 23e3 / 1e6:.0f} MB/GPU + scratch) lives in HBM/L2 and is accessed in place"}

    def system_file(self, tmp):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from synthetic import write_raster_system
        return write_raster_system(os.path.join(tmp, "raster_12x14.json"), 12, 14)

    def options(self, tmp=None):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from synthetic import UNIFORM_OPTIONS
        o = base_options()
        o.update(UNIFORM_OPTIONS)
        o.update({"max_total_staples": 168, "max_type_staples": 2, "staple_M": self.data["staple_M"], "simulation_type": "constant_temp",
                  "temp": self.data["bench_temp"], "random_seed": 20261020})
        o.pop("order_parameter_file")
        if tmp is not None:
            o["origami_input_filename"] = self.system_file(tmp)
        return o


WORKLOADS = {"ptmc": PTMC, "ct_four": CTFour, "ptmwus": PTMWUS, "anneal_large": AnnealLarge}


# ---------------------------------------------------------------------------------------------
# Reference arm / CPU baseline: the unmodified reference CLI (oracle/_ref), one process per host core
# ---------------------------------------------------------------------------------------------

def reference_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "latticeDNAOrigami"))


def reference_sample(wl, cores, moves_per_proc, workdir, step_k):
    """One bounded sample: `cores` independent processes of the unmodified reference CLI, each restarted from one of
    the workload's committed states and running `moves_per_proc` attempted moves (the reference's exchange traffic
    is 28 doubles per 100 moves, so real MPI would not change throughput; none is installed). Returns wall seconds."""
    cli = os.path.join(ROOT, "oracle", "_ref", "latticeDNAOrigami")
    inps = []
    for c in range(cores):
        d = os.path.join(workdir, f"s{step_k}_{c}")
        os.makedirs(d, exist_ok=True)
        opts, state = wl.reference_job(c, step_k, cores)
        opts = dict(opts)
        window = opts.pop("_window", None)
        if window is not None:
            bias = json.load(open(os.path.join(INPUTS, "biases_mwus-numfulldomains.json")))
            bias["origami"]["bias_functions"] = [b for b in bias["origami"]["bias_functions"] if b["type"] != "Grid"]
            bias["origami"]["bias_functions"][0].update({"min_op": window[0], "max_op": window[1]})
            json.dump(bias, open(os.path.join(d, "bias.json"), "w"))
            opts["bias_functions_file"] = os.path.join(d, "bias.json")
        if isinstance(wl, AnnealLarge):
            opts["origami_input_filename"] = wl.system_file(d)
        opts.update({"simulation_type": "constant_temp", "ct_steps": moves_per_proc, "random_seed": 1000 * (step_k + 1) + c,
                     "max_duration": 1e9, "output_filebase": os.path.join(d, "out"), "logging_freq": 0,
                     "restart_traj_file": write_trj(os.path.join(d, "start.trj"), state), "restart_step": 0})
        inps.append(write_inp(os.path.join(d, "ref.inp"), opts))
    t0 = time.perf_counter()
    running = [subprocess.Popen([cli, "-i", inp], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for inp in inps]
    for p in running:
        if p.wait() != 0:
            raise RuntimeError("reference process failed")
    return time.perf_counter() - t0


def ref_moves_per_proc(wl):
    return {"ct_four": 8 * REF_MOVES_PER_PROC, "anneal_large": REF_MOVES_PER_PROC // 8}.get(wl.name, REF_MOVES_PER_PROC)


def sample_text(wl, cores, moves):
    return (f"{cores} processes of the unmodified reference CLI (oracle/_ref), one per host core, {moves} constant-T moves each, "
            f"restarted from the workload's committed stationary states (.trj restart), every process at its state's "
            f"temperature / window")


def cpu_baseline(wl):
    cores = os.cpu_count() or 1
    moves = ref_moves_per_proc(wl)
    with tempfile.TemporaryDirectory(prefix="ldo_ref_") as tmp:
        wall = reference_sample(wl, cores, moves, tmp, 0)
    return {"value": cores * moves / wall, "unit": "attempted MC moves/s", "cores": cores, "kind": "reference",
            "sample": sample_text(wl, cores, moves) + f"; {wall:.1f} s wall"}


def run_reference_arm(args, rank, wl):
    if rank != 0:
        return
    if not reference_available():
        emit(({"impl": "reference", "unavailable": "oracle/_ref/latticeDNAOrigami not built (needs /root/reference at build time)"}))
        return
    cores = os.cpu_count() or 1
    moves = ref_moves_per_proc(wl)
    with tempfile.TemporaryDirectory(prefix="ldo_ref_") as tmp:
        for w in range(args.warmup):
            reference_sample(wl, cores, max(moves // 40, 100), tmp, 1000 + w)
        t0 = time.perf_counter()
        for k in range(args.steps):
            reference_sample(wl, cores, moves, tmp, k)
        wall = time.perf_counter() - t0
    value = cores * moves * args.steps / wall
    sample = "per step: " + sample_text(wl, cores, moves)
    emit(({
        "impl": "reference", "metric": wl.metric, "value": value,
        "unit": "attempted MC moves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": wl.config(args.gpus, args.replicas_per_gpu or wl.default_replicas),
        "cpu_baseline": {"value": value, "unit": "attempted MC moves/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "attempted MC moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------
# This repo's arm
# ---------------------------------------------------------------------------------------------

class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        if shutil.which("nvidia-smi") is None:
            return
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                      "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.proc.terminate()
        self.proc.wait()
        sm, smax, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v == "Active":
                    reasons.add(name)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


class DevArray:
    """__cuda_array_interface__ view of an engine-owned device buffer (for torch.as_tensor)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def tile_states(eng, states_of_local, n_units, unit, seed, subsequence_of):
    """Load `unit` starting states into replicas 0..unit-1 (ldo_set_state), replicate their checkpoint blobs over the
    other n_units - 1 groups of `unit` replicas (blobs are self-contained), then give every replica its own Philox
    subsequence."""
    import numpy as np
    for b, st in enumerate(states_of_local):
        eng.set_state(b, st["chain_index"], st["chain_ident"], st["chain_len"], np.asarray(st["pos"]).reshape(-1, 3),
                      np.asarray(st["ore"]).reshape(-1, 3))
    blob = eng.checkpoint_save(first=0, count=unit)
    eng.checkpoint_load(np.tile(blob, n_units), first=0, count=n_units * unit)  # one upload, one unpack launch
    eng.synchronize()
    eng.seed_subsequences(seed, [subsequence_of(r) for r in range(n_units * unit)])


def regime_rate(Simulation, tmp, system, temp, R, warm, launches, moves):
    """Attempted moves/s of a constant-T ensemble (CUDA events around `launches` run launches of `moves` moves)."""
    import torch
    o = base_options()
    o.update({"origami_input_filename": os.path.join(INPUTS, system), "simulation_type": "constant_temp", "temp": temp, "random_seed": 99})
    sim = Simulation(write_inp(os.path.join(tmp, f"regime{int(temp)}.inp"), o), R, torch.cuda.current_device())
    eng = sim.engine
    stream = torch.cuda.ExternalStream(eng.stream())
    eng.run(warm)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(launches):
        eng.run_async(moves)
    b.record(stream)
    eng.synchronize()
    eng.assert_ok()
    rate = R * launches * moves / (a.elapsed_time(b) * 1e-3)
    sim.close()
    return rate


def run_ours(args, rank, world, local_rank, wl):
    import numpy as np
    import torch

    from latticednaorigami_b200.binding import Simulation

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    R = args.replicas_per_gpu or wl.default_replicas
    tmp = tempfile.mkdtemp(prefix="ldo_bench_")
    is_pt = isinstance(wl, PTMC)
    cf, ccf = 100000, 1000000
    if is_pt:
        L = len(LADDER)
        slots = L // world
        opts = wl.options()
        sim = Simulation(write_inp(os.path.join(tmp, f"bench{rank}.inp"), opts), R, local_rank, rank=rank, n_ranks=world)
        eng = sim.engine
        # local slot b of this rank holds ladder slot k (serpentine dealing, ldo_exchange_pt)
        slot_of = [b * world + (world - 1 - rank if b & 1 else rank) for b in range(slots)]
        tile_states(eng, [wl.data["slots"][k] for k in slot_of], R // slots, slots, opts["random_seed"],
                    lambda r: (r // slots) * L + slot_of[r % slots] + (R // slots) * L * 0)
    else:
        if isinstance(wl, (PTMWUS, AnnealLarge)):
            opts = wl.options(tmp)
        else:
            opts = wl.options()
        unit = len(WINDOWS) if isinstance(wl, PTMWUS) else len(wl.data["states"])
        R -= R % unit
        sim = Simulation(write_inp(os.path.join(tmp, f"bench{rank}.inp"), opts), R, local_rank, rank=rank, n_ranks=world)
        eng = sim.engine
        if isinstance(wl, PTMWUS):
            per_window = {}
            for st in wl.data["states"]:
                per_window.setdefault(tuple(st["window"]), st)
            starts = [per_window[w] for w in WINDOWS]
        else:
            starts = wl.data["states"]
        tile_states(eng, starts, R // unit, unit, opts["random_seed"], lambda r: rank * R + r)
    eng.assert_ok()
    stream = torch.cuda.ExternalStream(eng.stream(), device=local_rank)
    if world > 1 and is_pt:
        # the exchange lives in the C++ host: one NCCL communicator across the ranks, created from a unique id that
        # rank 0 makes and torch.distributed only carries to the others
        from latticednaorigami_b200.binding import comm_unique_id
        box = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        sim.comm_init(box[0])

    swap = [0]
    kernel_events = []
    if isinstance(wl, PTMWUS):
        n_lad = R // len(WINDOWS)
        w2r = np.tile(np.arange(len(WINDOWS), dtype=np.int32), n_lad)
        w_att = np.zeros(n_lad * (len(WINDOWS) - 1), dtype=np.int64)
        w_acc = np.zeros_like(w_att)

    def one_step(record=False):
        swap[0] += 1
        if record:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
        if is_pt:
            # ldo_sim_exchange_round: moves, collection of the exchange records, ncclAllGather (N > 1), swap decisions
            # and the energy rebuild, all enqueued on the engine's stream by the C++ host - no host synchronisation
            sim.exchange_round(swap[0])
        else:
            eng.run_async(wl.moves_per_step, cf, 0, ccf)
        if record:
            b.record(stream)
            kernel_events.append((a, b))
        if is_pt:
            pass
        elif isinstance(wl, PTMWUS):
            eng.exchange_windows(swap[0], n_lad, len(WINDOWS), 1, [0], w2r, w_att, w_acc)

    def barrier():
        eng.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        one_step()
    barrier()
    eng.assert_ok()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launch_count()
    att0, acc0 = eng.move_stats()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record(stream)
    for _ in range(args.steps):
        one_step(record=True)
    end.record(stream)
    barrier()
    elapsed_ms = start.elapsed_time(end)
    launches = eng.launch_count() - launches0
    clocks = sampler.stop()
    eng.assert_ok()
    kernel_ms = sum(a.elapsed_time(b) for a, b in kernel_events) / len(kernel_events)
    step_ms = [a.elapsed_time(b) for a, b in kernel_events]

    if world > 1:
        t = torch.tensor([elapsed_ms], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    moves_per_step = R * world * wl.moves_per_step
    value = moves_per_step * args.steps / (elapsed_ms * 1e-3)
    att, acc = eng.move_stats()
    accepted_frac = float((acc - acc0).sum()) / float((att - att0).sum())
    staples = float(eng.counters()[:, 0].mean())

    # end to end through the C-ABI with host buffers: every step uploads the replicas' checkpoint blobs
    # from pinned host memory, runs the step, and reads blobs + energies back
    blob_bytes = eng.checkpoint_size() * R
    host_blob = torch.empty(blob_bytes, dtype=torch.uint8, pin_memory=True)
    eng.checkpoint_save(host_blob)
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.checkpoint_load(host_blob)
        one_step()
        eng.checkpoint_save(host_blob)
        energies = eng.energies()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = moves_per_step * e2e_steps / e2e_s
    eng.assert_ok()

    regimes = None
    if rank == 0 and world == 1 and is_pt and not args.no_regimes:
        # the two ends of the ladder on their own: constant-T ensembles of the same size
        regimes = {"unbound_345K": regime_rate(Simulation, tmp, "snodin_unbound.json", 345, R, 300, 3, 100),
                   "assembled_330K": regime_rate(Simulation, tmp, "snodin_assembled.json", 330, R, 100, 2, 100),
                   "unit": "attempted MC moves/s", "replicas": R}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
        if "hbm_gbs" in peaks:
            peak, peak_src = peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"
        state_bytes = eng.state_bytes()
        hbm_bytes = 2.0 * state_bytes * R  # algorithmic: every replica's state is loaded and stored once per launch
        achieved = hbm_bytes / (kernel_ms * 1e-3) / 1e9
        moves_per_launch = R * wl.moves_per_step
        staged = state_bytes < 8192
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": ("k_exec_staged<CapsSmall>" if staged else "k_exec_inplace<CapsLarge>") + f" (run, {wl.moves_per_step} moves/replica)",
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": hbm_bytes, "peak_source": peak_src,
                "kernel_ms_note": "CUDA events on the engine's stream around one step" +
                                  (" (run launch + the exchange's collection / decision / energy-rebuild launches, < 1 % of it: "
                                   "profiles/launches_r2.csv)" if is_pt else "")}
        prof_path = os.path.join(ROOT, "profiles", f"ncu_{wl.name}_traffic.json")
        if os.path.exists(prof_path):
            t = json.load(open(prof_path))
            if t.get("replicas") == R:
                roof["traffic"] = t["traffic_bytes_per_launch"]
                roof["issue"] = dict(t.get("issue") or {}, peak_ipc_per_sm=4.0)
                roof["smem_device_count"] = t.get("smem")
        if staged and is_pt:
            smem_gbs = SMEM_BYTES_PER_MOVE * moves_per_launch / (kernel_ms * 1e-3) / 1e9
            if "smem_gbs" in peaks:
                smem_peak, smem_src = peaks["smem_gbs"], "measured (MEASURED_PEAKS.json smem_gbs)"
            else:
                sp = os.path.join(ROOT, "profiles", "smem_peak_r2.json")
                if os.path.exists(sp):
                    smem_peak, smem_src = json.load(open(sp))["smem_gbs"], "measured (profiles/smem_peak_r2.json, profiles/smem_bw.cu)"
                else:
                    smem_peak, smem_src = 148 * 128 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e9, "nominal 148 SM * 128 B/clk * sm_max_mhz"
            roof["smem"] = {"achieved": smem_gbs, "peak": smem_peak, "unit": "GB/s", "frac": smem_gbs / smem_peak,
                            "bytes_per_move": SMEM_BYTES_PER_MOVE, "peak_source": smem_src}
            roof["note"] = ("the path is bound by instruction supply, not by HBM or shared memory (SURVEY.md 8d, profiles/README.md); "
                            "roofline.issue carries the ncu issue-slot figures of the same launch")
        line = {
            "metric": wl.metric, "value": value, "unit": "attempted MC moves/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": wl.config(world, R), "clocks": clocks, "gpu_launches": int(launches),
            "accepted_moves_per_s": value * accepted_frac, "mean_staples_per_replica": staples,
            "kernel_ms_first_last": [step_ms[0], step_ms[-1]],
            "e2e": {"value": e2e_value, "unit": "attempted MC moves/s", "h2d_bytes_per_step": int(blob_bytes),
                    "d2h_bytes_per_step": int(blob_bytes + energies.nbytes), "steps": e2e_steps},
            "roofline": roof,
        }
        if regimes is not None:
            line["regimes"] = regimes
        if world == 1 and reference_available() and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl)
        else:
            line["cpu_baseline"] = None
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    shutil.rmtree(tmp, ignore_errors=True)


def main():
    capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ptmc", choices=sorted(WORKLOADS))
    ap.add_argument("--replicas-per-gpu", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-regimes", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]()
    if args.impl == "reference":
        run_reference_arm(args, rank, wl)
        return
    if world != args.gpus:
        if args.gpus != 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    run_ours(args, rank, world, local_rank, wl)


if __name__ == "__main__":
    main()
