#!/bin/bash
# GPU call 18: single-site merges of batch three without the multi-site inlining; twist-test inline twin; against batch two.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_synthetic_systems.py -m "gpu and not slow" -x -q) > $O/c18_pytest.log 2>&1
echo "pytest rc=$?" >> $O/c18_pytest.log
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c18_bench.json 2> $O/c18_bench.err
LDO_B200_LIB=ab/lib_merge2.so timeout 300 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c18_bench_merge2.json 2> $O/c18_bench_merge2.err
LDO_B200_LIB=ab/lib_twist.so timeout 300 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c18_bench_twist.json 2> $O/c18_bench_twist.err
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c18_bench_again.json 2> $O/c18_bench_again.err
tail -3 $O/c18_pytest.log; for f in c18_bench c18_bench_merge2 c18_bench_twist c18_bench_again; do cut -c1-160 $O/$f.json; tail -1 $O/$f.err; done
