#!/bin/bash
# GPU call 23: the two single-site merges of the third batch one by one (experiment twins) against the final build.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
timeout 300 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c23_bench.json 2> $O/c23_bench.err
for v in xbind xwalks; do
  LDO_B200_LIB=ab/lib_$v.so timeout 300 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c23_bench_$v.json 2> $O/c23_bench_$v.err
done
timeout 300 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c23_bench_again.json 2> $O/c23_bench_again.err
for f in c23_bench c23_bench_xbind c23_bench_xwalks c23_bench_again; do cut -c1-160 $O/$f.json; tail -1 $O/$f.err; done
