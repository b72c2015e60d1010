"""Which call sites of System::step a move executes, from a host-emulation build with -DLDO_CALLER_PROFILE
(return-address histogram resolved with addr2line). Usage: python profiles/step_callers.py [system temp movetype_index]"""
import collections
import ctypes
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: E402
from latticednaorigami_b200 import binding  # noqa: E402
from latticednaorigami_b200.binding import Simulation  # noqa: E402

system = sys.argv[1] if len(sys.argv) > 1 else "snodin_assembled.json"
temp = float(sys.argv[2]) if len(sys.argv) > 2 else 330
which = int(sys.argv[3]) if len(sys.argv) > 3 else 3
tmp = tempfile.mkdtemp()
so = os.path.join(tmp, "libhs_prof.so")
csrc = os.path.join(ROOT, "latticednaorigami_b200", "csrc")
subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-DLDO_HOSTSIM", "-DLDO_CALLER_PROFILE", "-shared", "-o", so, "-x", "c++",
                os.path.join(csrc, "ldo_engine.cu"), os.path.join(csrc, "ldo_host.cpp"), os.path.join(csrc, "ldo_sim.cpp")], check=True, capture_output=True)
raw = ctypes.CDLL(so)
lib = binding.bind(raw)
ms = json.load(open(os.path.join(conftest.INPUTS, "moveset_standard.json")))["origami"]["movetypes"]
p = os.path.join(tmp, "ms.json")
json.dump({"origami": {"movetypes": [dict(ms[which], freq="1")]}}, open(p, "w"))
o = conftest.make_options(system, temp=temp, random_seed=3)
o["movetype_file"] = p
sim = Simulation(conftest.write_inp(os.path.join(tmp, "a.inp"), o), 1, 0, lib=lib)
sim.engine.run(100)
callers = (ctypes.c_ulonglong * 8192).in_dll(raw, "ldo_dbg_callers")
for i in range(8192):
    callers[i] = 0
n = 2000
sim.engine.run(n)
base = None
for line in open("/proc/self/maps"):
    if so in line:
        base = int(line.split("-")[0], 16)
        break
hist = {callers[2 * i] - base: callers[2 * i + 1] for i in range(4096) if callers[2 * i]}
addrs = sorted(hist, key=lambda a: -hist[a])
out = subprocess.run(["addr2line", "-f", "-C", "-i", "-e", so] + [hex(a - 1) for a in addrs], capture_output=True, text=True).stdout.split("\n")
# addr2line -i prints (function, file:line) pairs, innermost first; regroup per address by re-running one at a time is
# slow, so resolve the frames individually
total = sum(hist.values())
print(f"{ms[which]['type']} on {system} at {temp} K: {total / n:.0f} step calls per move")
by_site = collections.Counter()
for a in addrs:
    r = subprocess.run(["addr2line", "-f", "-C", "-i", "-e", so, hex(a - 1)], capture_output=True, text=True).stdout.strip().split("\n")
    frames = []
    for k in range(0, len(r), 2):
        fn = r[k].split("(")[0].split("::")[-1]
        frames.append(f"{fn}@{r[k + 1].split('/')[-1].split(' ')[0]}")
    by_site[" < ".join(frames[:3])] += hist[a]
for site, c in by_site.most_common(40):
    print(f"{c / n:8.1f} {100 * c / total:5.1f}%  {site}")
