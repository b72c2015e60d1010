#!/bin/bash
# GPU call 22: recoil growth and the old-configuration pass merged into rg_regrow_and_test (their second call sites are in the
# linker move) - A/B against the final build.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c22_bench.json 2> $O/c22_bench.err
LDO_B200_LIB=ab/lib_merge2.so timeout 300 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c22_bench_merge2.json 2> $O/c22_bench_merge2.err
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c22_bench_again.json 2> $O/c22_bench_again.err
for f in c22_bench c22_bench_merge2 c22_bench_again; do cut -c1-160 $O/$f.json; tail -1 $O/$f.err; done
