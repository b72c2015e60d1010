"""Summarises an ncu report of the run kernel: headline raw metrics, stall mix and a per-function table
(instructions / samples / stall reasons) built from the SASS page and the cubin symbol table.
    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep [library of the profiled build] > profiles/ncu_<name>.txt
The symbol table must come from the build that was profiled (default: the library in the tree).
"""
import csv, subprocess, collections, bisect, sys
rep = sys.argv[1]
import os
libso = os.path.abspath(sys.argv[2]) if len(sys.argv) > 2 else os.path.abspath('latticednaorigami_b200/libldo_b200.so')
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_elapsed','launch__shared_mem_per_block_dynamic','launch__waves_per_multiprocessor','smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
for i,h in enumerate(hdr):
    if h in want: print(f"{h:70s} {units[i]:16s} {vals[i]}")
sass = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows = list(csv.reader(sass.splitlines()))
hdr = rows[1]; idx = {h:i for i,h in enumerate(hdr)}
syms=[]
out = subprocess.run("D=$(mktemp -d) && cd $D && cuobjdump -xelf all " + libso + " >/dev/null 2>&1; readelf -sW *.sm_100a.cubin 2>/dev/null | awk '$4==\"FUNC\" || $4==\"NOTYPE\" {print $2, $3, $8}' | grep 'k_exec_stagedIN3ldo4CapsILi80E'", shell=True, capture_output=True, text=True).stdout
for line in out.splitlines():
    v,size,name = line.split()[:3]; syms.append((int(v,16), name))
syms.sort(); starts=[s[0] for s in syms]
def short(n):
    n = n.split('$')[-1]
    d = subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
    import re as _re
    d = _re.sub(r'ldo::Caps<[^>]*>', 'K', d)
    return d.split('(')[0][-58:]
base = int(rows[2][0],16)
inst=collections.Counter(); samp=collections.Counter(); st=collections.defaultdict(collections.Counter)
stall_cols=[(h,i) for h,i in idx.items() if h.startswith('stall_') and 'Not Issued' not in h]
tot=collections.Counter()
for r in rows[2:]:
    if len(r)<len(hdr): continue
    off=int(r[0],16)-base; k=bisect.bisect_right(starts,off)-1
    name=syms[k][1] if k>=0 else '?'
    inst[name]+=int(r[idx['Instructions Executed']]); samp[name]+=int(r[idx['# Samples']])
    for h,i in stall_cols:
        v=int(r[i]); st[name][h]+=v; tot[h]+=v
ti=sum(inst.values()); ts=sum(samp.values())
print('total inst', ti, 'samples', ts, 'sass instrs', len(rows)-2)
print({k:round(100*v/ts,1) for k,v in tot.most_common(7)})
print(f"{'function':60s} inst%  samp%  no_inst long_sb wait  short_sb branch")
for n,v in samp.most_common(28):
    s=st[n]; sm=max(1,samp[n])
    print(f"{short(n):60s} {100*inst[n]/ti:5.1f} {100*v/ts:6.1f} {100*s['stall_no_inst']/sm:6.1f} {100*s['stall_long_sb']/sm:6.1f} {100*s['stall_wait']/sm:6.1f} {100*s['stall_short_sb']/sm:6.1f} {100*s['stall_branch_resolving']/sm:6.1f}")
