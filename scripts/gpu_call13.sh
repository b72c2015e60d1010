#!/bin/bash
# GPU call 13: chain neighbours computed once per FourBody evaluation - parity on the device and same-box A/B.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_synthetic_systems.py tests/test_enumeration.py "tests/test_production_parity.py::test_snodin_production_matches_reference_mc" -m "gpu and not slow" -x -q) > $O/c13_pytest.log 2>&1
echo "pytest rc=$?" >> $O/c13_pytest.log
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c13_bench.json 2> $O/c13_bench.err
LDO_B200_LIB=ab/lib_prevsteps.so timeout 300 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c13_bench_prev.json 2> $O/c13_bench_prev.err
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c13_bench_again.json 2> $O/c13_bench_again.err
LDO_B200_LIB=ab/lib_prevsteps.so timeout 300 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c13_bench_prev2.json 2> $O/c13_bench_prev2.err
tail -3 $O/c13_pytest.log; for f in c13_bench c13_bench_prev c13_bench_again c13_bench_prev2; do cut -c1-160 $O/$f.json; tail -1 $O/$f.err; done
