"""ORACLE — test infrastructure only (imported by tests/, __graft_entry__.smoke and bench.py's
cpu_baseline / --impl reference legs; never by the product package).

ctypes wrapper around ``oracle/_ref/liboracle_ref.so``: the UNMODIFIED reference sources of
cumberworth/LatticeDNAOrigami compiled by ``oracle/Makefile`` plus the C-ABI driver
``oracle/src/ref_driver.cpp``. See that file for the reference file:line each call follows.
"""
import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "liboracle_ref.so")
CLI_PATH = os.path.join(HERE, "_ref", "latticeDNAOrigami")

DRAW_DTYPE = np.dtype(
    [("kind", "<i4"), ("lo", "<i4"), ("hi", "<i4"), ("ival", "<i4"), ("real", "<f8")]
)

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        # RTLD_DEEPBIND: the reference objects must bind to their own libstdc++ symbols even when
        # numpy/torch preloaded other C++ runtimes into the process.
        L = C.CDLL(LIB_PATH, mode=os.RTLD_NOW | os.RTLD_DEEPBIND)
        L.oref_create.restype = C.c_void_p
        L.oref_create.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        L.oref_destroy.argtypes = [C.c_void_p]
        L.oref_last_error.restype = C.c_char_p
        L.oref_last_error.argtypes = [C.c_void_p]
        L.oref_simulate.argtypes = [C.c_void_p, C.c_longlong]
        L.oref_step.restype = C.c_longlong
        L.oref_step.argtypes = [C.c_void_p]
        L.oref_seed.argtypes = [C.c_void_p, C.c_int]
        L.oref_tape_len.restype = C.c_longlong
        L.oref_tape_len.argtypes = [C.c_void_p]
        L.oref_tape_copy.argtypes = [C.c_void_p, C.c_void_p]
        L.oref_tape_clear.argtypes = [C.c_void_p]
        L.oref_tape_set_replay.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
        L.oref_num_movetypes.argtypes = [C.c_void_p]
        L.oref_move_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oref_num_chains.argtypes = [C.c_void_p]
        L.oref_num_domains.argtypes = [C.c_void_p]
        L.oref_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.oref_set_state.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.oref_energy.restype = C.c_double
        L.oref_energy.argtypes = [C.c_void_p]
        L.oref_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.oref_energy_split.argtypes = [C.c_void_p, C.c_void_p]
        L.oref_check_all_constraints.argtypes = [C.c_void_p]
        L.oref_center.argtypes = [C.c_void_p, C.c_int]
        L.oref_update_temp.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.oref_update_staple_us.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.oref_update_staple_us.restype = None
        L.oref_num_staple_types.argtypes = [C.c_void_p]
        L.oref_staple_us.argtypes = [C.c_void_p, C.c_void_p]
        L.oref_pair_energies.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.oref_init_energies.argtypes = [C.c_void_p, C.c_void_p]
        L.oref_order_param.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.oref_order_param_stored.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]
        L.oref_total_bias.restype = C.c_double
        L.oref_total_bias.argtypes = [C.c_void_p]
        L.oref_total_bias_stored.restype = C.c_double
        L.oref_total_bias_stored.argtypes = [C.c_void_p]
        L.oref_check_domain.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oref_set_domain.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oref_unassign_domain.restype = C.c_double
        L.oref_unassign_domain.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oref_nn_unitless_thermo.argtypes = [C.c_char_p, C.c_double, C.c_double, C.c_void_p]
        L.oref_nn_unitless_energy.restype = C.c_double
        L.oref_nn_unitless_energy.argtypes = [C.c_char_p, C.c_double, C.c_double]
        L.oref_nn_longest_contig_complement.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        L.oref_num_walks.restype = C.c_double
        L.oref_num_walks.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.oref_pt_run.restype = C.c_void_p
        L.oref_pt_run.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.oref_pt_destroy.argtypes = [C.c_void_p]
        L.oref_pt_tape_len.restype = C.c_longlong
        L.oref_pt_tape_len.argtypes = [C.c_void_p, C.c_int]
        L.oref_pt_tape_copy.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oref_pt_num_flags.restype = C.c_longlong
        L.oref_pt_num_flags.argtypes = [C.c_void_p, C.c_int]
        L.oref_pt_flags.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oref_pt_num_marks.restype = C.c_longlong
        L.oref_pt_num_marks.argtypes = [C.c_void_p]
        L.oref_pt_marks.argtypes = [C.c_void_p, C.c_void_p]
        L.oref_pt_num_chains.argtypes = [C.c_void_p, C.c_int]
        L.oref_pt_num_domains.argtypes = [C.c_void_p, C.c_int]
        L.oref_pt_get_state.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.oref_pt_energy.restype = C.c_double
        L.oref_pt_energy.argtypes = [C.c_void_p, C.c_int]
        L.oref_pt_num_movetypes.argtypes = [C.c_void_p, C.c_int]
        L.oref_pt_move_stats.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oref_pt_acceptance_p.argtypes = [C.c_void_p] + [C.c_double] * 8 + [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5
        _lib = L
    return _lib


def write_inp(path, options):
    """Write a reference-format key=value parameter file (parser.cpp:477-479)."""
    with open(path, "w") as f:
        for k, v in options.items():
            if isinstance(v, bool):
                v = "true" if v else "false"
            elif isinstance(v, (list, tuple)):
                v = " ".join(str(x) for x in v)
            f.write(f"{k}={v}\n")


# Defaults of examples/constant-temp.inp with all outputs off
SNODIN_OPTIONS = {
    "domain_type": "HalfTurn",
    "binding_pot": "FourBody",
    "misbinding_pot": "Opposing",
    "stacking_pot": "Constant",
    "hybridization_pot": "NearestNeighbour",
    "apply_mean_field_cor": False,
    "temp": 330,
    "staple_M": 1e-7,
    "cation_M": 0.5,
    "staple_u_mult": 1,
    "stacking_ene": -1000,
    "max_total_staples": 24,
    "max_type_staples": 12,
    "max_staple_size": 2,
    "domain_update_biases_present": False,
    "simulation_type": "constant_temp",
    "centering_freq": 0,
    "constraint_check_freq": 0,
    "max_duration": 1e9,
    "ct_steps": 0,
    "logging_freq": 0,
    "configs_output_freq": 0,
    "vtf_output_freq": 0,
    "vcf_per_domain": False,
    "counts_output_freq": 0,
    "order_params_output_freq": 0,
    "times_output_freq": 0,
    "energies_output_freq": 0,
}


class RefSystem:
    """One reference OrigamiSystem (+ ConstantTGCMCSimulation when with_sim)."""

    def __init__(self, options, with_sim=True, workdir=None):
        self._tmp = None
        if workdir is None:
            self._tmp = tempfile.TemporaryDirectory(prefix="oref_")
            workdir = self._tmp.name
        opts = dict(options)
        opts.setdefault("output_filebase", os.path.join(workdir, "out"))
        inp = os.path.join(workdir, "oracle.inp")
        write_inp(inp, opts)
        err = C.create_string_buffer(1024)
        self.L = lib()
        self.h = self.L.oref_create(inp.encode(), 1 if with_sim else 0, err, 1024)
        if not self.h:
            raise RuntimeError("oracle: " + err.value.decode())

    def close(self):
        if self.h:
            self.L.oref_destroy(self.h)
            self.h = None
        if self._tmp is not None:
            self._tmp.cleanup()
            self._tmp = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + self.L.oref_last_error(self.h).decode())

    # stepping
    def seed(self, seed):
        self.L.oref_seed(self.h, int(seed))

    def simulate(self, steps):
        self._check(self.L.oref_simulate(self.h, int(steps)))

    @property
    def step(self):
        return self.L.oref_step(self.h)

    def tape(self, clear=False):
        n = self.L.oref_tape_len(self.h)
        out = np.zeros(n, dtype=DRAW_DTYPE)
        if n:
            self.L.oref_tape_copy(self.h, out.ctypes.data)
        if clear:
            self.L.oref_tape_clear(self.h)
        return out

    def set_replay(self, tape):
        tape = np.ascontiguousarray(tape, dtype=DRAW_DTYPE)
        self.L.oref_tape_set_replay(self.h, tape.ctypes.data, len(tape))

    def move_stats(self):
        n = self.L.oref_num_movetypes(self.h)
        a = np.zeros(n, dtype=np.int64)
        b = np.zeros(n, dtype=np.int64)
        self.L.oref_move_stats(self.h, a.ctypes.data, b.ctypes.data)
        return a, b

    # state
    def state(self):
        nc = self.L.oref_num_chains(self.h)
        nd = self.L.oref_num_domains(self.h)
        ci = np.zeros(nc, dtype=np.int32)
        cid = np.zeros(nc, dtype=np.int32)
        cl = np.zeros(nc, dtype=np.int32)
        pos = np.zeros((nd, 3), dtype=np.int32)
        ore = np.zeros((nd, 3), dtype=np.int32)
        st = np.zeros(nd, dtype=np.int32)
        bd = np.zeros((nd, 2), dtype=np.int32)
        self.L.oref_get_state(
            self.h, ci.ctypes.data, cid.ctypes.data, cl.ctypes.data, pos.ctypes.data,
            ore.ctypes.data, st.ctypes.data, bd.ctypes.data)
        return {"chain_index": ci, "chain_ident": cid, "chain_len": cl, "pos": pos,
                "ore": ore, "state": st, "bound": bd}

    def set_state(self, chain_index, chain_ident, chain_len, pos, ore):
        ci = np.ascontiguousarray(chain_index, dtype=np.int32)
        cid = np.ascontiguousarray(chain_ident, dtype=np.int32)
        cl = np.ascontiguousarray(chain_len, dtype=np.int32)
        p = np.ascontiguousarray(pos, dtype=np.int32)
        o = np.ascontiguousarray(ore, dtype=np.int32)
        self._check(self.L.oref_set_state(
            self.h, len(ci), ci.ctypes.data, cid.ctypes.data, cl.ctypes.data,
            p.ctypes.data, o.ctypes.data))

    def energy(self):
        return self.L.oref_energy(self.h)

    def counters(self):
        out = np.zeros(9, dtype=np.int32)
        self.L.oref_counters(self.h, out.ctypes.data)
        keys = ["staples", "domains", "bound_pairs", "fully_bound_pairs", "self_bound_pairs",
                "misbound_pairs", "stacked_pairs", "unassigned", "current_c_i"]
        return dict(zip(keys, (int(x) for x in out)))

    def energy_split(self):
        out = np.zeros(3)
        self.L.oref_energy_split(self.h, out.ctypes.data)
        return {"enthalpy": out[0], "entropy": out[1], "stacking": out[2]}

    def check_all_constraints(self):
        self._check(self.L.oref_check_all_constraints(self.h))

    def center(self, d=0):
        self.L.oref_center(self.h, d)

    def update_temp(self, temp, stacking_mult=1.0):
        self._check(self.L.oref_update_temp(self.h, temp, stacking_mult))

    def staple_us(self, temp, staple_u_mult=1.0):
        """m_staple_us after OrigamiSystem::update_staple_us(temp, mult) (origami_system.cpp:624-628)."""
        self.L.oref_update_staple_us(self.h, temp, staple_u_mult)
        n = self.L.oref_num_staple_types(self.h)
        out = np.zeros(max(n, 1))
        self.L.oref_staple_us(self.h, out.ctypes.data)
        return out[:n]

    def pair_energies(self, a, b):
        out = np.zeros(4)
        if self.L.oref_pair_energies(self.h, a, b, out.ctypes.data) != 0:
            return None
        return out

    def init_energies(self):
        out = np.zeros(3)
        self.L.oref_init_energies(self.h, out.ctypes.data)
        return out

    def order_param(self, tag):
        v = C.c_int(0)
        self._check(self.L.oref_order_param(self.h, tag.encode(), C.byref(v)))
        return v.value

    def order_param_stored(self, tag):
        """(m_param, m_defined) as they stand, without the calc_param that order_param() does."""
        v, d = C.c_int(0), C.c_int(0)
        self._check(self.L.oref_order_param_stored(self.h, tag.encode(), C.byref(v), C.byref(d)))
        return v.value, bool(d.value)

    def total_bias(self):
        return self.L.oref_total_bias(self.h)

    def total_bias_stored(self):
        """SystemBiases::get_total_bias as it stands (what .ene prints), without the re-evaluation total_bias() does."""
        return self.L.oref_total_bias_stored(self.h)

    def check_domain(self, c, d, pos, ore):
        p = np.asarray(pos, dtype=np.int32)
        o = np.asarray(ore, dtype=np.int32)
        out = np.zeros(1)
        v = self.L.oref_check_domain(self.h, c, d, p.ctypes.data, o.ctypes.data, out.ctypes.data)
        return bool(v), out[0]

    def set_domain(self, c, d, pos, ore):
        p = np.asarray(pos, dtype=np.int32)
        o = np.asarray(ore, dtype=np.int32)
        out = np.zeros(1)
        v = self.L.oref_set_domain(self.h, c, d, p.ctypes.data, o.ctypes.data, out.ctypes.data)
        return bool(v), out[0]

    def unassign_domain(self, c, d):
        return self.L.oref_unassign_domain(self.h, c, d)


def pt_run(options, n_ranks, seeds, record_tapes=True, workdir=None):
    """The reference's own replica-exchange driver (PTGCMCSimulation::run, ptmc_simulation.cpp:106-150), one thread
    per rank over the thread-backed boost::mpi shim. `options` must name a *_parallel_tempering simulation_type and
    an output_filebase (the reference always writes <filebase>-<rank>.out/.vsf and <filebase>.swp). Returns a dict:
    per-rank tapes, final states, energies, move statistics; rank 0's exchange draws separated from its MC draws
    (`exchange_reals`, `mc_tapes[0]`); and the rows of the .swp file."""
    tmp = None
    if workdir is None:
        tmp = tempfile.TemporaryDirectory(prefix="oref_pt_")
        workdir = tmp.name
    opts = dict(options)
    opts.setdefault("output_filebase", os.path.join(workdir, "pt"))
    inp = os.path.join(workdir, "oracle_pt.inp")
    write_inp(inp, opts)
    L = lib()
    err = C.create_string_buffer(1024)
    sd = np.ascontiguousarray(seeds, dtype=np.int32)
    h = L.oref_pt_run(inp.encode(), n_ranks, sd.ctypes.data, 1 if record_tapes else 0, err, 1024)
    if not h:
        raise RuntimeError("oracle: " + err.value.decode())
    out = {"tapes": [], "states": [], "energies": [], "attempts": [], "accepts": []}
    for r in range(n_ranks):
        n = L.oref_pt_tape_len(h, r)
        t = np.zeros(n, dtype=DRAW_DTYPE)
        if n:
            L.oref_pt_tape_copy(h, r, t.ctypes.data)
        out["tapes"].append(t)
        nc, nd = L.oref_pt_num_chains(h, r), L.oref_pt_num_domains(h, r)
        ci, cid, cl = (np.zeros(nc, dtype=np.int32) for _ in range(3))
        pos, ore = np.zeros((nd, 3), dtype=np.int32), np.zeros((nd, 3), dtype=np.int32)
        L.oref_pt_get_state(h, r, ci.ctypes.data, cid.ctypes.data, cl.ctypes.data, pos.ctypes.data, ore.ctypes.data)
        out["states"].append({"chain_index": ci, "chain_ident": cid, "chain_len": cl, "pos": pos, "ore": ore})
        out["energies"].append(L.oref_pt_energy(h, r))
        nm = L.oref_pt_num_movetypes(h, r)
        a, b = np.zeros(nm, dtype=np.int64), np.zeros(nm, dtype=np.int64)
        L.oref_pt_move_stats(h, r, a.ctypes.data, b.ctypes.data)
        out["attempts"].append(a)
        out["accepts"].append(b)
    # umbrella-sampling drivers: draws of the multi-window driver's own generator (rank 0: the exchange tests)
    flags = []
    for r in range(n_ranks):
        nf = L.oref_pt_num_flags(h, r)
        f = np.zeros(nf, dtype=np.uint8)
        if nf:
            L.oref_pt_flags(h, r, f.ctypes.data)
        flags.append(f.astype(bool))
    nm = L.oref_pt_num_marks(h)
    marks = np.zeros(nm, dtype=np.int64)
    if nm:
        L.oref_pt_marks(h, marks.ctypes.data)
    L.oref_pt_destroy(h)
    marks = marks.reshape(-1, 2)
    out["marks"] = marks
    # rank 0: MC draws with the master's exchange draws cut out, and those draws on their own
    t0 = out["tapes"][0]
    keep = np.ones(len(t0), dtype=bool)
    for b, e in marks:
        keep[b:e] = False
    if len(flags[0]) == len(t0) and len(t0):
        keep = ~flags[0]
    out["mc_tapes"] = [t0[keep]] + out["tapes"][1:]
    ex = t0[~keep]
    assert np.all(ex["kind"] == 0)
    out["exchange_reals"] = ex["real"].copy()
    # where every exchange round's draws start in exchange_reals
    out["exchange_offsets"] = np.concatenate([[0], np.cumsum(marks[:, 1] - marks[:, 0])]).astype(np.int64)
    swp = opts["output_filebase"] + ".swp"
    if os.path.exists(swp):
        rows = open(swp).read().splitlines()
        out["swp_header"] = rows[0]
        out["swp"] = [[int(x) for x in row.split()] for row in rows[1:]]
    if tmp is not None:
        tmp.cleanup()
    return out


def us_run(options, n_ranks, seeds, workdir):
    """The reference's umbrella-sampling drivers (us_simulation.cpp): umbrella_sampling (one rank),
    mw_umbrella_sampling / ptmw_umbrella_sampling (one rank = one thread per window). Output files (.biases per
    iteration, ...) are left in `workdir` under options['output_filebase']; returns what pt_run returns."""
    return pt_run(options, n_ranks, seeds, record_tapes=True, workdir=workdir)


def pt_acceptance_p(ref, temps, umults, bmults, smults, dep1, dep2, staple_u1, staple_u2, staple_n1, staple_n2):
    """PTGCMCSimulation::calc_acceptance_p of the unmodified reference (ptmc_simulation.cpp:275-313) for one pair
    of replicas; `ref` must have been created from ut_parallel_tempering options. The per-staple arrays carry
    n_types entries (the reference's loop reads that many, App. A1)."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (dep1, dep2, staple_u1, staple_u2, staple_n1, staple_n2)]
    out = np.zeros(1)
    rc = lib().oref_pt_acceptance_p(ref.h, temps[0], temps[1], umults[0], umults[1], bmults[0], bmults[1], smults[0], smults[1],
                                    arrs[0].ctypes.data, arrs[1].ctypes.data, len(arrs[2]), arrs[2].ctypes.data,
                                    arrs[3].ctypes.data, arrs[4].ctypes.data, arrs[5].ctypes.data, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(lib().oref_last_error(ref.h).decode())
    return float(out[0])


def nn_unitless_thermo(seq, temp, cation_M):
    out = np.zeros(2)
    lib().oref_nn_unitless_thermo(seq.encode(), temp, cation_M, out.ctypes.data)
    return out[0], out[1]


def nn_unitless_energy(seq, temp, cation_M):
    return lib().oref_nn_unitless_energy(seq.encode(), temp, cation_M)


def nn_longest_contig_complement(a, b):
    buf = C.create_string_buffer(4096)
    lib().oref_nn_longest_contig_complement(a.encode(), b.encode(), buf, 4096)
    return [s for s in buf.value.decode().split("\n") if s]


def num_walks(start, end, steps):
    s = np.asarray(start, dtype=np.int32)
    e = np.asarray(end, dtype=np.int32)
    return lib().oref_num_walks(s.ctypes.data, e.ctypes.data, steps)
