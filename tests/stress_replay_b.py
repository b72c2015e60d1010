"""Second replay campaign against the live oracle (CPU, host emulation of the device sources): the cases stress_replay.py
does not visit - 168-domain rasters (in-place capacities) at two assembly regimes, a cyclic 24-domain raster, misbinding
Disallowed, the mean-field correction, four_unbound, snodin with distance order parameters and well biases - over the
standard, CTCB and linker movesets; state bit-exact, tape fully consumed, running energy to 1e-12 of the run's energy scale
after every chunk. Not collected by pytest (needs oracle/_ref):  python tests/stress_replay_b.py SECONDS
Round 2, final build: 5720 runs in 780 s, 0 failures."""
import sys, os, tempfile, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, ROOT)
import conftest, oracle_ref as o
from latticednaorigami_b200.binding import Simulation
from synthetic import UNIFORM_OPTIONS, write_raster_system
tmp = tempfile.mkdtemp()
t_end = time.time() + float(sys.argv[1])
def run(opts, seed, chunks, steps):
    r = o.RefSystem(opts); r.seed(seed)
    sim = Simulation(conftest.write_inp(os.path.join(tmp, f"{seed}.inp"), opts), 1, 0, lib=conftest.load_hostsim())
    scale = 1.0
    for k in range(chunks):
        r.tape(clear=True); r.simulate(steps); tape = r.tape(clear=True)
        sim.engine.attach_tape(0, tape); sim.engine.run(steps, opts.get("centering_freq", 0), 0, opts.get("constraint_check_freq", 0))
        st, info = sim.engine.status()
        if st[0] != 0: return f"status {st[0]} {info[0]} chunk {k}"
        try:
            conftest.assert_state_equal(sim.engine.state(0), r.state(), "x")
            assert sim.engine.tape_position(0) == len(tape), "tape"
            e = r.energy(); scale = max(scale, abs(e)); assert abs(sim.engine.energies()[0, 0] - e) <= 1e-12 * scale, "energy"
        except AssertionError as ex: return f"MISMATCH {ex} chunk {k}"
    return "ok"
I = conftest.INPUTS
seed = 50000; n = 0; bad = 0
def raster(w, h, cyc, temp, mt, M, ms, **kw):
    sysf = write_raster_system(os.path.join(tmp, f"r{w}x{h}_{int(cyc)}.json"), w, h, cyc)
    opts = conftest.make_options(temp=temp, max_total_staples=mt, max_type_staples=2, staple_M=M, **UNIFORM_OPTIONS)
    opts["origami_input_filename"] = sysf; opts["movetype_file"] = os.path.join(I, ms); opts.update(kw)
    return opts
while time.time() < t_end:
    cases = []
    for ms in ("moveset_standard.json", "moveset_ctcb.json", "moveset_linker.json"):
        cases.append(("large", raster(12, 14, False, 270, 168, 1.0, ms, centering_freq=50, constraint_check_freq=40), 4, 100))
        cases.append(("large295", raster(12, 14, False, 295, 168, 1e-3, ms), 4, 100))
        cases.append(("mid cyc", raster(4, 6, True, 290, 24, 1e-2, ms), 6, 100))
        cases.append(("disallowed", raster(3, 4, False, 300, 8, 1e-4, ms, misbinding_pot="Disallowed"), 6, 100))
        cases.append(("meanfield", raster(3, 4, False, 300, 8, 1e-4, ms, apply_mean_field_cor=True), 6, 100))
        for temp in (325, 345):
            cases.append(("four", conftest.make_options("four_unbound.json", "moveset_four.json", temp=temp, max_total_staples=2, max_type_staples=2), 6, 100))
        b = conftest.make_options("snodin_assembled.json", ms, temp=336)
        b["order_parameter_file"] = os.path.join(I, "ops_dist.json"); b["bias_functions_file"] = os.path.join(I, "biases_dist.json")
        cases.append(("dist biases", b, 6, 50))
    for name, opts, chunks, steps in cases:
        if time.time() > t_end: break
        seed += 1
        try: res = run(opts, seed, chunks, steps)
        except Exception as ex: res = f"EXC {type(ex).__name__} {str(ex)[:120]}"
        n += 1
        if res != "ok": bad += 1; print(name, opts["movetype_file"].split("/")[-1], seed, res, flush=True)
print("runs", n, "failures", bad)
