"""Turns one ncu capture of the run launch (+ the instruction-cache / divergence counters of the same launch) into the
JSON bench.py reads for roofline.traffic / roofline.issue / roofline.smem.
    python profiles/make_traffic_json.py gpurun_out/c6_run100.ncu-rep gpurun_out/c6_icc.csv profiles/ncu_r2_final.txt > profiles/ncu_ptmc_traffic.json
"""
import csv
import json
import subprocess
import sys

rep, icc_csv, summary = sys.argv[1], sys.argv[2], sys.argv[3]
replicas, moves = 16384, 100
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]


def metric(name):
    i = hdr.index(name)
    v = float(vals[i].replace(",", ""))
    u = units[i]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1.0)
    return v * scale


icc = {}
lines = open(icc_csv).read().splitlines()
lines = lines[next(i for i, l in enumerate(lines) if l.startswith('"ID"')):]
for r in csv.DictReader(lines):
    if r.get("Metric Name"):
        icc[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
stalls = {}
for line in open(summary):
    if line.startswith("{'stall_"):
        stalls = eval(line)
rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
inst = metric("smsp__inst_executed.sum")
wf = metric("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
out = {
    "source": f"{summary} (ncu --set full --clock-control none of profiles/prof_run.py --moves {moves} --replicas {replicas}: the run "
              "launch bench.py times, stationary ladder)",
    "replicas": replicas, "moves_per_launch": moves,
    "traffic_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
    "kernel_ms_under_ncu": metric("gpu__time_duration.sum") / 1e6 if units[hdr.index("gpu__time_duration.sum")] in ("ns", "nsecond") else metric("gpu__time_duration.sum"),
    "issue": {
        "warp_instructions": inst, "warp_instructions_per_move": inst / (replicas * moves),
        "ipc_per_sm": metric("sm__inst_executed.avg.per_cycle_elapsed"),
        "issue_active_pct": metric("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "stall_no_inst_pct": stalls.get("stall_no_inst"),
        "stall_branch_resolving_pct": stalls.get("stall_branch_resolving"),
        "warps_active_per_scheduler": metric("smsp__warps_active.avg.per_cycle_active"),
        "warps_eligible_per_scheduler": metric("smsp__warps_eligible.avg.per_cycle_active"),
        "thread_inst_per_warp_inst": icc.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "icc_lookups": icc.get("sm__icc_requests.sum"),
        "icc_hit_pct": 100.0 * icc.get("sm__icc_requests_lookup_hit.sum", 0) / max(1.0, icc.get("sm__icc_requests.sum", 1)),
    },
    "smem": {"wavefronts": wf, "wavefronts_per_move": wf / (replicas * moves), "bytes_per_move_at_128B": 128 * wf / (replicas * moves)},
}
print(json.dumps(out, indent=1))
