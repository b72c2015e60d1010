#!/bin/bash
# GPU call 7: device enumeration (tests, timing), the bench lines of the final build (default with cpu_baseline, reference
# arm, the other workloads).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
(time timeout 900 python -m pytest tests/test_enumeration.py tests/test_cli.py -m gpu -x -q -s) > $O/c7_enum.log 2>&1
echo "pytest rc=$?" >> $O/c7_enum.log
timeout 900 python bench.py > $O/c7_bench.json 2> $O/c7_bench.err
timeout 900 python bench.py --impl reference > $O/c7_bench_ref.json 2> $O/c7_bench_ref.err
for w in ct_four ptmwus anneal_large; do
  timeout 600 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline > $O/c7_bench_$w.json 2> $O/c7_bench_$w.err
done
tail -12 $O/c7_enum.log; for f in c7_bench c7_bench_ref c7_bench_ct_four c7_bench_ptmwus c7_bench_anneal_large; do cut -c1-220 $O/$f.json; tail -2 $O/$f.err; done
