#!/bin/bash
# GPU call 20: stacked flags of the four adjacent pairs computed once per FourBody evaluation - parity and A/B against the build before.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_synthetic_systems.py -m "gpu and not slow" -x -q) > $O/c20_pytest.log 2>&1
echo "pytest rc=$?" >> $O/c20_pytest.log
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c20_bench.json 2> $O/c20_bench.err
LDO_B200_LIB=ab/lib_merge2.so timeout 300 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c20_bench_merge2.json 2> $O/c20_bench_merge2.err
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c20_bench_again.json 2> $O/c20_bench_again.err
tail -3 $O/c20_pytest.log; for f in c20_bench c20_bench_merge2 c20_bench_again; do cut -c1-160 $O/$f.json; tail -1 $O/$f.err; done
