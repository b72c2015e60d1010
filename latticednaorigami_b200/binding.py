"""ctypes binding of the C-ABI in include/ldo_b200.h and include/ldo_host.h."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libldo_b200.so")

DRAW_DTYPE = np.dtype([("kind", "<i4"), ("lo", "<i4"), ("hi", "<i4"), ("ival", "<i4"), ("real", "<f8")])

# Symbols declared in include/ldo_b200.h and include/ldo_host.h (checked by tests/test_abi.py)
ENGINE_SYMBOLS = [
    "ldo_engine_create", "ldo_engine_destroy", "ldo_last_error", "ldo_num_replicas",
    "ldo_set_temperature_tables", "ldo_set_moveset", "ldo_set_domain_update_biases", "ldo_set_order_params", "ldo_set_biases",
    "ldo_set_window", "ldo_set_grid_bias", "ldo_get_grid_visits", "ldo_set_control", "ldo_get_control",
    "ldo_seed", "ldo_seed_subsequences", "ldo_rng_state_words", "ldo_get_rng_state", "ldo_set_rng_state", "ldo_attach_tape", "ldo_tape_position", "ldo_set_state", "ldo_state_capacity",
    "ldo_get_state", "ldo_run", "ldo_get_status", "ldo_run_async", "ldo_synchronize", "ldo_stream",
    "ldo_get_energies", "ldo_get_counters", "ldo_get_staple_counts", "ldo_get_order_params",
    "ldo_get_move_stats", "ldo_get_run_timing", "ldo_recompute_energies", "ldo_check_all_constraints", "ldo_center",
    "ldo_set_exchange_ladder", "ldo_exchange_collect", "ldo_exchange_pt", "ldo_exchange_pt_2d", "ldo_exchange_acceptance_p", "ldo_get_reduced_staple_u", "ldo_set_step", "ldo_exchange_buffers",
    "ldo_exchange_windows", "ldo_exchange_collect_async", "ldo_exchange_state_set", "ldo_exchange_state_get", "ldo_exchange_pt_async", "ldo_set_exchange_tape", "ldo_exchange_tape_status", "ldo_set_reference_draw_order", "ldo_build_info", "ldo_launch_count", "ldo_state_bytes", "ldo_checkpoint_size", "ldo_checkpoint_save", "ldo_checkpoint_load",
    "ldo_enumerate_conformations", "ldo_get_exchange_mults", "ldo_replace_config", "ldo_enable_move_trackers", "ldo_get_move_trackers", "ldo_get_linker_trackers",
]
HOST_SYMBOLS = [
    "ldo_host_last_error", "ldo_comm_unique_id", "ldo_sim_comm_init", "ldo_sim_exchange_round", "ldo_sim_create", "ldo_sim_destroy", "ldo_sim_engine", "ldo_sim_run",
    "ldo_sim_exchange_advance", "ldo_sim_exchange_apply", "ldo_sim_exchange_state", "ldo_sim_num_temps",
    "ldo_sim_num_order_params", "ldo_sim_order_param_tag", "ldo_sim_num_movetypes",
    "ldo_sim_movetype_label", "ldo_sim_num_staple_types", "ldo_sim_step", "ldo_sim_pair_energies",
    "ldo_sim_init_energies", "ldo_host_nn_unitless_thermo", "ldo_host_longest_contig_complement",
    "ldo_host_no_walks", "ldo_host_energy_tables", "ldo_host_inp_value", "ldo_sim_enumeration_summary",
]

STATUS_NAMES = {
    0: "ok", 1: "tape exhausted", 2: "tape mismatch", 3: "coordinate out of range", 4: "occupancy table full",
    5: "move scratch capacity exceeded", 6: "unassigned domain at constraint check", 7: "stack count inconsistency",
    8: "system energy inconsistency", 9: "constraints in violation", 10: "contiguous domains not adjacent",
    11: "binding to an already bound domain", 12: "setting an already assigned domain",
    13: "nonsensical exchange probability", 14: "system has unbound staple", 15: "internal error",
}


class LdoError(RuntimeError):
    pass


_libs = {}


def load(path=None):
    """Load the CUDA library. Fails loudly when it is missing or is not a CUDA build: there is no CPU path.
    `path` / LDO_B200_LIB may name another CUDA build of the same sources (profiling variants under ab/)."""
    path = path or os.environ.get("LDO_B200_LIB") or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise LdoError(
            f"{path} not found: build the CUDA library first (make, or __graft_entry__.build()); "
            "latticednaorigami_b200 has no CPU fallback")
    L = bind(C.CDLL(path))
    info = L.ldo_build_info().decode()
    if not info.startswith("cuda"):
        raise LdoError(f"{path} is not a CUDA build of the engine ({info!r}); latticednaorigami_b200 has no CPU path")
    _libs[path] = L
    return L


def bind(L):
    """Attach the C-ABI signatures of include/ldo_b200.h and include/ldo_host.h to an already opened library."""
    if not hasattr(L, "ldo_build_info"):
        raise LdoError("library exports no ldo_build_info marker")
    vp, i, d, ll = C.c_void_p, C.c_int, C.c_double, C.c_longlong
    sig = {
        "ldo_engine_create": (i, [vp, i, i, vp]),
        "ldo_engine_destroy": (None, [vp]),
        "ldo_last_error": (C.c_char_p, [vp]),
        "ldo_num_replicas": (i, [vp]),
        "ldo_set_temperature_tables": (i, [vp, i, i, vp, vp, vp, vp, vp]),
        "ldo_set_moveset": (i, [vp, i, vp, i]),
        "ldo_set_domain_update_biases": (i, [vp, i]),
        "ldo_set_order_params": (i, [vp, i, vp]),
        "ldo_set_biases": (i, [vp, i, vp]),
        "ldo_set_window": (i, [vp, i, i, i, i]),
        "ldo_set_grid_bias": (i, [vp, i, i, vp, vp, vp]),
        "ldo_get_grid_visits": (i, [vp, i, i, vp, i]),
        "ldo_set_control": (i, [vp, i, i, vp, vp, vp, vp]),
        "ldo_get_control": (i, [vp, i, i, vp, vp, vp, vp]),
        "ldo_seed": (i, [vp, C.c_ulonglong, C.c_uint]),
        "ldo_seed_subsequences": (i, [vp, C.c_ulonglong, vp]),
        "ldo_rng_state_words": (i, []),
        "ldo_get_rng_state": (i, [vp, i, i, vp]),
        "ldo_set_rng_state": (i, [vp, i, i, vp]),
        "ldo_attach_tape": (i, [vp, i, vp, ll]),
        "ldo_tape_position": (i, [vp, i, vp]),
        "ldo_set_state": (i, [vp, i, i, vp, vp, vp, vp, vp]),
        "ldo_state_capacity": (i, [vp, vp, vp]),
        "ldo_get_state": (i, [vp, i, vp, vp, vp, vp, vp, vp, vp, vp]),
        "ldo_run": (i, [vp, ll, i, i, i]),
        "ldo_run_async": (i, [vp, ll, i, i, i]),
        "ldo_synchronize": (i, [vp]),
        "ldo_stream": (vp, [vp]),
        "ldo_get_status": (i, [vp, vp, vp]),
        "ldo_get_energies": (i, [vp, vp]),
        "ldo_get_counters": (i, [vp, vp]),
        "ldo_get_staple_counts": (i, [vp, vp]),
        "ldo_get_order_params": (i, [vp, vp]),
        "ldo_get_move_stats": (i, [vp, vp, vp]),
        "ldo_get_run_timing": (i, [vp, vp]),
        "ldo_recompute_energies": (i, [vp, vp, vp]),
        "ldo_check_all_constraints": (i, [vp]),
        "ldo_center": (i, [vp, i]),
        "ldo_set_exchange_ladder": (i, [vp, i, vp, vp, vp, vp]),
        "ldo_exchange_collect": (i, [vp, vp]),
        "ldo_exchange_pt": (i, [vp, i, ll, i, i, i, i, vp, vp, vp, vp]),
        "ldo_exchange_pt_2d": (i, [vp, ll, i, i, i, i, i, vp, vp, vp, vp]),
        "ldo_get_reduced_staple_u": (i, [vp, vp]),
        "ldo_set_step": (i, [vp, ll]),
        "ldo_exchange_acceptance_p": (C.c_double, [i, vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, vp, vp]),
        "ldo_exchange_buffers": (i, [vp, i, vp, vp, vp]),
        "ldo_exchange_windows": (i, [vp, ll, i, i, i, i, vp, vp, vp, vp]),
        "ldo_set_reference_draw_order": (i, [vp, i]),
        "ldo_exchange_collect_async": (i, [vp]),
        "ldo_exchange_state_set": (i, [vp, i, i, vp, vp, vp]),
        "ldo_exchange_state_get": (i, [vp, i, i, vp, vp, vp]),
        "ldo_exchange_pt_async": (i, [vp, i, i, ll, i, i, i, i]),
        "ldo_set_exchange_tape": (i, [vp, vp, ll, vp, ll]),
        "ldo_exchange_tape_status": (i, [vp, vp, vp]),
        "ldo_build_info": (C.c_char_p, []),
        "ldo_launch_count": (ll, [vp]),
        "ldo_state_bytes": (C.c_ulong, [vp]),
        "ldo_checkpoint_size": (C.c_ulong, [vp]),
        "ldo_checkpoint_save": (i, [vp, i, i, vp]),
        "ldo_checkpoint_load": (i, [vp, i, i, vp]),
        "ldo_enumerate_conformations": (i, [vp, vp, i, vp, vp, vp, vp, vp]),
        "ldo_get_exchange_mults": (i, [vp, i, vp]),
        "ldo_replace_config": (i, [vp, i, i, vp, vp, vp, vp, vp]),
        "ldo_enable_move_trackers": (i, [vp, i]),
        "ldo_get_move_trackers": (i, [vp, i, vp, vp]),
        "ldo_get_linker_trackers": (i, [vp, i, vp, vp, vp, vp]),
        "ldo_sim_enumeration_summary": (i, [vp, vp]),
        "ldo_host_last_error": (C.c_char_p, []),
        "ldo_sim_create": (vp, [C.c_char_p, i, i, i, i]),
        "ldo_sim_destroy": (None, [vp]),
        "ldo_sim_engine": (vp, [vp]),
        "ldo_sim_run": (i, [vp]),
        "ldo_sim_exchange_advance": (i, [vp]),
        "ldo_sim_exchange_round": (i, [vp, ll]),
        "ldo_comm_unique_id": (i, [vp]),
        "ldo_sim_comm_init": (i, [vp, vp]),
        "ldo_sim_exchange_apply": (i, [vp, ll, vp]),
        "ldo_sim_exchange_state": (i, [vp, vp, vp, vp]),
        "ldo_sim_num_temps": (i, [vp]),
        "ldo_sim_num_order_params": (i, [vp]),
        "ldo_sim_order_param_tag": (C.c_char_p, [vp, i]),
        "ldo_sim_num_movetypes": (i, [vp]),
        "ldo_sim_movetype_label": (C.c_char_p, [vp, i]),
        "ldo_sim_num_staple_types": (i, [vp]),
        "ldo_sim_step": (ll, [vp]),
        "ldo_sim_pair_energies": (i, [vp, i, i, i, vp]),
        "ldo_sim_init_energies": (i, [vp, i, vp]),
        "ldo_host_nn_unitless_thermo": (i, [C.c_char_p, d, d, vp]),
        "ldo_host_longest_contig_complement": (i, [C.c_char_p, C.c_char_p, C.c_char_p, i]),
        "ldo_host_no_walks": (i, [vp, vp, i]),
        "ldo_host_energy_tables": (i, [C.c_char_p, d, vp, vp, vp, vp, vp, vp]),
        "ldo_host_inp_value": (i, [C.c_char_p, C.c_char_p, C.c_char_p, i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    return L


def _ptr(a):
    return a.ctypes.data if a is not None else None


class Engine:
    """View of an ``ldo_engine*`` (owned by a :class:`Simulation` or created raw by tests)."""

    def __init__(self, lib, handle, n_ops=0, n_movetypes=0, n_staple_types=0):
        self.L = lib
        self.h = handle
        self.R = lib.ldo_num_replicas(handle)
        self.n_ops = n_ops
        self.n_movetypes = n_movetypes
        self.n_staple_types = n_staple_types

    def _check(self, rc):
        if rc != 0:
            raise LdoError(self.L.ldo_last_error(self.h).decode())

    # stepping
    def run(self, n_steps, centering_freq=0, centering_domain=0, constraint_check_freq=0):
        self._check(self.L.ldo_run(self.h, int(n_steps), centering_freq, centering_domain, constraint_check_freq))

    def run_async(self, n_steps, centering_freq=0, centering_domain=0, constraint_check_freq=0):
        self._check(self.L.ldo_run_async(self.h, int(n_steps), centering_freq, centering_domain, constraint_check_freq))

    def synchronize(self):
        self._check(self.L.ldo_synchronize(self.h))

    def stream(self):
        return self.L.ldo_stream(self.h)

    def status(self):
        st = np.zeros(self.R, dtype=np.int32)
        dt = np.zeros(self.R, dtype=np.int32)
        self._check(self.L.ldo_get_status(self.h, _ptr(st), _ptr(dt)))
        return st, dt

    def assert_ok(self):
        st, dt = self.status()
        bad = np.nonzero(st)[0]
        if len(bad):
            r = int(bad[0])
            raise LdoError(f"replica {r}: {STATUS_NAMES.get(int(st[r]), st[r])} (detail {int(dt[r])})")

    def set_reference_draw_order(self, on=True):
        self._check(self.L.ldo_set_reference_draw_order(self.h, 1 if on else 0))

    def seed(self, seed, first_subsequence=0):
        self._check(self.L.ldo_seed(self.h, int(seed), int(first_subsequence)))

    def seed_subsequences(self, seed, subsequences):
        sub = np.ascontiguousarray(subsequences, dtype=np.uint32)
        assert len(sub) == self.R
        self._check(self.L.ldo_seed_subsequences(self.h, int(seed), _ptr(sub)))

    def rng_state(self, first=0, count=None):
        count = self.R - first if count is None else count
        w = np.zeros((count, self.L.ldo_rng_state_words()), dtype=np.uint64)
        self._check(self.L.ldo_get_rng_state(self.h, first, count, _ptr(w)))
        return w

    def set_rng_state(self, words, first=0):
        w = np.ascontiguousarray(words, dtype=np.uint64).reshape(-1, self.L.ldo_rng_state_words())
        self._check(self.L.ldo_set_rng_state(self.h, first, len(w), _ptr(w)))

    def attach_tape(self, replica, tape):
        tape = np.ascontiguousarray(tape, dtype=DRAW_DTYPE)
        self._check(self.L.ldo_attach_tape(self.h, replica, _ptr(tape), len(tape)))

    def tape_position(self, replica):
        v = C.c_longlong(0)
        self._check(self.L.ldo_tape_position(self.h, replica, C.byref(v)))
        return v.value

    # configuration
    def set_state(self, replica, chain_index, chain_ident, chain_len, pos, ore):
        ci = np.ascontiguousarray(chain_index, dtype=np.int32)
        cid = np.ascontiguousarray(chain_ident, dtype=np.int32)
        cl = np.ascontiguousarray(chain_len, dtype=np.int32)
        p = np.ascontiguousarray(pos, dtype=np.int32)
        o = np.ascontiguousarray(ore, dtype=np.int32)
        self._check(self.L.ldo_set_state(self.h, replica, len(ci), _ptr(ci), _ptr(cid), _ptr(cl), _ptr(p), _ptr(o)))

    def state(self, replica):
        mc, md = C.c_int(0), C.c_int(0)
        self.L.ldo_state_capacity(self.h, C.byref(mc), C.byref(md))
        nc = C.c_int(0)
        ci = np.zeros(mc.value, dtype=np.int32)
        cid = np.zeros(mc.value, dtype=np.int32)
        cl = np.zeros(mc.value, dtype=np.int32)
        pos = np.zeros((md.value, 3), dtype=np.int32)
        ore = np.zeros((md.value, 3), dtype=np.int32)
        st = np.zeros(md.value, dtype=np.int32)
        bd = np.zeros((md.value, 2), dtype=np.int32)
        self._check(self.L.ldo_get_state(self.h, replica, C.byref(nc), _ptr(ci), _ptr(cid), _ptr(cl), _ptr(pos), _ptr(ore), _ptr(st), _ptr(bd)))
        n = nc.value
        nd = int(cl[:n].sum())
        return {"chain_index": ci[:n], "chain_ident": cid[:n], "chain_len": cl[:n], "pos": pos[:nd],
                "ore": ore[:nd], "state": st[:nd], "bound": bd[:nd]}

    # control
    def set_control(self, first, temp_idx=None, staple_u_mult=None, bias_mult=None, stacking_mult=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=t)
                for a, t in ((temp_idx, np.int32), (staple_u_mult, np.float64), (bias_mult, np.float64), (stacking_mult, np.float64))]
        count = max(len(a) for a in arrs if a is not None)
        self._check(self.L.ldo_set_control(self.h, first, count, *[_ptr(a) for a in arrs]))

    def control(self):
        ti = np.zeros(self.R, dtype=np.int32)
        a = np.zeros(self.R)
        b = np.zeros(self.R)
        c = np.zeros(self.R)
        self._check(self.L.ldo_get_control(self.h, 0, self.R, _ptr(ti), _ptr(a), _ptr(b), _ptr(c)))
        return {"temp_idx": ti, "staple_u_mult": a, "bias_mult": b, "stacking_mult": c}

    # observables
    def energies(self):
        out = np.zeros((self.R, 5))
        self._check(self.L.ldo_get_energies(self.h, _ptr(out)))
        return out

    def counters(self):
        out = np.zeros((self.R, 9), dtype=np.int32)
        self._check(self.L.ldo_get_counters(self.h, _ptr(out)))
        return out

    def staple_counts(self):
        out = np.zeros((self.R, max(self.n_staple_types, 1)), dtype=np.int32)
        self._check(self.L.ldo_get_staple_counts(self.h, _ptr(out)))
        return out

    def order_params(self):
        out = np.zeros((self.R, max(self.n_ops, 1)), dtype=np.int32)
        self._check(self.L.ldo_get_order_params(self.h, _ptr(out)))
        return out

    def run_timing(self):
        """[R][3]: start ns, end ns, SM id of every replica's warp in the last run launch (diagnostics)."""
        out = np.zeros((self.R, 3), dtype=np.int64)
        self._check(self.L.ldo_get_run_timing(self.h, _ptr(out)))
        return out

    def move_stats(self):
        a = np.zeros((self.R, max(self.n_movetypes, 1)), dtype=np.int64)
        b = np.zeros((self.R, max(self.n_movetypes, 1)), dtype=np.int64)
        self._check(self.L.ldo_get_move_stats(self.h, _ptr(a), _ptr(b)))
        return a, b

    def exchange_mults(self, movetype):
        """Exchange multipliers of a staple-exchange movetype, [R, n_staple_types] (adaptive ones are per replica)."""
        out = np.zeros((self.R, self.n_staple_types))
        self._check(self.L.ldo_get_exchange_mults(self.h, int(movetype), _ptr(out)))
        return out

    def enable_move_trackers(self, on=True):
        self._check(self.L.ldo_enable_move_trackers(self.h, int(on)))

    def move_trackers(self, replica):
        """(sticky[n_movetypes, 2], counts[n_movetypes, 2 fields, 64 values, (attempts, accepts)])"""
        sticky = np.zeros((self.n_movetypes, 2), dtype=np.int32)
        counts = np.zeros((self.n_movetypes, 2, 64, 2), dtype=np.uint32)
        self._check(self.L.ldo_get_move_trackers(self.h, int(replica), _ptr(sticky), _ptr(counts)))
        return sticky, counts

    def linker_trackers(self, replica):
        """(sticky[n_movetypes, 6], entries[n, (movetype, table, value a, value b, attempts, accepts)], dropped)"""
        sticky = np.zeros((self.n_movetypes, 6), dtype=np.int32)
        entries = np.zeros((1024, 6), dtype=np.int32)
        n, dropped = C.c_int(0), C.c_int(0)
        self._check(self.L.ldo_get_linker_trackers(self.h, int(replica), _ptr(sticky), C.byref(n), _ptr(entries), C.byref(dropped)))
        return sticky, entries[:n.value].copy(), dropped.value

    def recompute_energies(self):
        e = np.zeros(self.R)
        s = np.zeros(self.R, dtype=np.int32)
        self._check(self.L.ldo_recompute_energies(self.h, _ptr(e), _ptr(s)))
        return e, s

    def check_all_constraints(self):
        self._check(self.L.ldo_check_all_constraints(self.h))

    def center(self, centering_domain=0):
        self._check(self.L.ldo_center(self.h, centering_domain))

    # biases
    def set_window(self, replica, bias, lo, hi):
        self._check(self.L.ldo_set_window(self.h, replica, bias, lo, hi))

    def set_grid_bias(self, replica, bias, lo, n, values):
        lo = np.ascontiguousarray(lo, dtype=np.int32)
        n = np.ascontiguousarray(n, dtype=np.int32)
        v = np.ascontiguousarray(values, dtype=np.float64)
        self._check(self.L.ldo_set_grid_bias(self.h, replica, bias, _ptr(lo), _ptr(n), _ptr(v)))

    def grid_visits(self, replica, bias, size, clear=False):
        out = np.zeros(size, dtype=np.int64)
        self._check(self.L.ldo_get_grid_visits(self.h, replica, bias, _ptr(out), 1 if clear else 0))
        return out

    def exchange_windows(self, swap_i, n_ladders, n_windows, grid_bias, window_biases, window_to_replica, attempts, accepts):
        wb = np.ascontiguousarray(window_biases, dtype=np.int32)
        self._check(self.L.ldo_exchange_windows(self.h, int(swap_i), n_ladders, n_windows, grid_bias, len(wb), _ptr(wb),
                                                _ptr(window_to_replica), _ptr(attempts), _ptr(accepts)))

    # checkpoint
    def launch_count(self):
        return self.L.ldo_launch_count(self.h)

    def checkpoint_size(self):
        return self.L.ldo_checkpoint_size(self.h)

    def checkpoint_save(self, out=None, first=0, count=None):
        count = self.R - first if count is None else count
        if out is None:
            out = np.zeros(count * self.checkpoint_size(), dtype=np.uint8)
        ptr = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
        self._check(self.L.ldo_checkpoint_save(self.h, first, count, ptr))
        return out

    def checkpoint_load(self, blob, first=0, count=None):
        count = self.R - first if count is None else count
        ptr = blob.data_ptr() if hasattr(blob, "data_ptr") else blob.ctypes.data
        self._check(self.L.ldo_checkpoint_load(self.h, first, count, ptr))

    # exchange
    def exchange_collect(self, to_host=True):
        if not to_host:
            self._check(self.L.ldo_exchange_collect(self.h, None))
            return None
        out = np.zeros((self.R, 4 + self.n_staple_types))
        self._check(self.L.ldo_exchange_collect(self.h, _ptr(out)))
        return out

    def set_exchange_tape(self, reals, round_offsets):
        r = np.ascontiguousarray(reals, dtype=np.float64)
        o = np.ascontiguousarray(round_offsets, dtype=np.int64)
        self._check(self.L.ldo_set_exchange_tape(self.h, _ptr(r), len(r), _ptr(o), len(o) - 1))

    def exchange_tape_status(self):
        missing, unused = C.c_longlong(0), C.c_longlong(0)
        self._check(self.L.ldo_exchange_tape_status(self.h, C.byref(missing), C.byref(unused)))
        return missing.value, unused.value

    def exchange_buffers(self, n_global):
        send, recv, nq = C.c_void_p(0), C.c_void_p(0), C.c_int(0)
        self._check(self.L.ldo_exchange_buffers(self.h, n_global, C.byref(send), C.byref(recv), C.byref(nq)))
        return send.value, recv.value, nq.value

    def state_bytes(self):
        return self.L.ldo_state_bytes(self.h)


class Simulation:
    """A simulation described by a reference-format ``.inp`` file (ldo_sim_create)."""

    def __init__(self, inp_path, n_replicas=1, device=0, rank=0, n_ranks=1, lib_path=None, lib=None):
        # `lib`: an already bound library handle (the test-suite's host emulation is loaded by tests/conftest.py,
        # never through load())
        self.L = lib if lib is not None else load(lib_path)
        self.h = self.L.ldo_sim_create(os.fsencode(inp_path), n_replicas, device, rank, n_ranks)
        if not self.h:
            raise LdoError(self.L.ldo_host_last_error().decode())
        self.n_ops = self.L.ldo_sim_num_order_params(self.h)
        self.n_movetypes = self.L.ldo_sim_num_movetypes(self.h)
        self.n_staple_types = self.L.ldo_sim_num_staple_types(self.h)
        self.engine = Engine(self.L, self.L.ldo_sim_engine(self.h), self.n_ops, self.n_movetypes, self.n_staple_types)
        self.op_tags = [self.L.ldo_sim_order_param_tag(self.h, i).decode() for i in range(self.n_ops)]
        self.movetype_labels = [self.L.ldo_sim_movetype_label(self.h, i).decode() for i in range(self.n_movetypes)]

    def close(self):
        if self.h:
            self.L.ldo_sim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise LdoError(self.L.ldo_host_last_error().decode())
        return rc

    def run(self):
        self._check(self.L.ldo_sim_run(self.h))

    @property
    def step(self):
        return self.L.ldo_sim_step(self.h)

    def enumeration_summary(self):
        """After run() with simulation_type=enumerate: configurations (with multiplicities), average energy, average
        bias, conformations visited."""
        out = np.zeros(4)
        self._check(self.L.ldo_sim_enumeration_summary(self.h, _ptr(out)))
        return {"num_configs": out[0], "average_energy": out[1], "average_bias": out[2], "leaves": int(out[3])}

    def exchange_round(self, swap_i):
        """One whole exchange round on the engine's stream (moves, collection, NCCL all-gather, decisions)."""
        return self._check(self.L.ldo_sim_exchange_round(self.h, int(swap_i)))

    def comm_init(self, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self.L.ldo_sim_comm_init(self.h, buf))

    def exchange_advance(self):
        return self._check(self.L.ldo_sim_exchange_advance(self.h))

    def exchange_apply(self, swap_i, dependent_all=None):
        dep = None if dependent_all is None else np.ascontiguousarray(dependent_all, dtype=np.float64)
        self._check(self.L.ldo_sim_exchange_apply(self.h, int(swap_i), _ptr(dep)))

    def exchange_state(self, n_ladders, num_reps, two_d=False):
        """slot -> replica map and the swap counters: [n_ladders][num_reps - 1] for the 1-D variants,
        [n_ladders][2][num_reps] (direction, slot) for 2d_parallel_tempering."""
        q = np.zeros((n_ladders, num_reps), dtype=np.int32)
        shape = (n_ladders, 2, num_reps) if two_d else (n_ladders, num_reps - 1)
        a = np.zeros(shape, dtype=np.int64)
        b = np.zeros(shape, dtype=np.int64)
        self.L.ldo_sim_exchange_state(self.h, _ptr(q), _ptr(a), _ptr(b))
        return q, a, b

    def pair_energies(self, temp_idx, a, b):
        out = np.zeros(3)
        if self.L.ldo_sim_pair_energies(self.h, temp_idx, a, b, _ptr(out)) != 0:
            return None
        return out

    def init_energies(self, temp_idx):
        out = np.zeros(3)
        self.L.ldo_sim_init_energies(self.h, temp_idx, _ptr(out))
        return out


def comm_unique_id(lib_path=None, lib=None):
    """128-byte NCCL unique id (rank 0 creates it; every rank passes it to Simulation.comm_init)."""
    L = lib if lib is not None else load(lib_path)
    buf = C.create_string_buffer(128)
    if L.ldo_comm_unique_id(buf) != 0:
        raise LdoError(L.ldo_host_last_error().decode())
    return buf.raw


# ---- GPU-free host helpers (table builder, parameter-file reader) ---------------------------------

def host_energy_tables(inp_path, temp, lib_path=None, lib=None):
    L = lib if lib is not None else load(lib_path)
    n = C.c_int(0)
    if L.ldo_host_energy_tables(os.fsencode(inp_path), temp, C.byref(n), None, None, None, None, None) != 0:
        raise LdoError(L.ldo_host_last_error().decode())
    sz = (2 * n.value + 1) ** 2
    e, h, s = np.zeros(sz), np.zeros(sz), np.zeros(sz)
    present = np.zeros(sz, dtype=np.int8)
    init = np.zeros(3)
    if L.ldo_host_energy_tables(os.fsencode(inp_path), temp, C.byref(n), _ptr(e), _ptr(h), _ptr(s), _ptr(present), _ptr(init)) != 0:
        raise LdoError(L.ldo_host_last_error().decode())
    return {"n_ident": n.value, "energy": e, "enthalpy": h, "entropy": s, "present": present, "init": init}


def host_inp_value(inp_path, key, lib_path=None, lib=None):
    L = lib if lib is not None else load(lib_path)
    buf = C.create_string_buffer(4096)
    if L.ldo_host_inp_value(os.fsencode(inp_path), key.encode(), buf, 4096) != 0:
        raise LdoError(L.ldo_host_last_error().decode())
    return buf.value.decode()


def host_nn_unitless_thermo(seq, temp, cation_M, lib_path=None, lib=None):
    L = lib if lib is not None else load(lib_path)
    out = np.zeros(2)
    if L.ldo_host_nn_unitless_thermo(seq.encode(), temp, cation_M, _ptr(out)) != 0:
        raise LdoError(L.ldo_host_last_error().decode())
    return out[0], out[1]


def host_longest_contig_complement(a, b, lib_path=None, lib=None):
    L = lib if lib is not None else load(lib_path)
    buf = C.create_string_buffer(4096)
    L.ldo_host_longest_contig_complement(a.encode(), b.encode(), buf, 4096)
    return [x for x in buf.value.decode().split("\n") if x]


def host_no_walks(start, end, steps, lib_path=None, lib=None):
    L = lib if lib is not None else load(lib_path)
    s = np.asarray(start, dtype=np.int32)
    e = np.asarray(end, dtype=np.int32)
    return bool(L.ldo_host_no_walks(_ptr(s), _ptr(e), steps))
