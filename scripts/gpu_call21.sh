#!/bin/bash
# GPU call 21 (final build): bench lines of the other workloads, enumeration timing.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
for w in ct_four ptmwus anneal_large; do
  timeout 400 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline > $O/c21_bench_$w.json 2> $O/c21_bench_$w.err
done
timeout 300 python profiles/enum_time.py > $O/c21_enum_time.txt 2>&1
for w in ct_four ptmwus anneal_large; do cut -c1-200 $O/c21_bench_$w.json; tail -1 $O/c21_bench_$w.err; done; cat $O/c21_enum_time.txt
