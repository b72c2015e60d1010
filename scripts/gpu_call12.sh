#!/bin/bash
# GPU call 12: typed trackers / adaptive exchange / umbrella-sampling restart on the device; A/B of the tracker hooks.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
(time timeout 900 python -m pytest tests/test_restart_and_outputs.py tests/test_hostsim_replay.py tests/test_us_oracle.py tests/test_gpu_parity.py -m "gpu and not slow" -x -q) > $O/c12_pytest.log 2>&1
echo "pytest rc=$?" >> $O/c12_pytest.log
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c12_bench.json 2> $O/c12_bench.err
LDO_B200_LIB=ab/lib_notrackers.so timeout 300 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c12_bench_notrackers.json 2> $O/c12_bench_notrackers.err
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c12_bench_again.json 2> $O/c12_bench_again.err
LDO_B200_LIB=ab/lib_notrackers.so timeout 300 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c12_bench_notrackers2.json 2> $O/c12_bench_notrackers2.err
tail -3 $O/c12_pytest.log; for f in c12_bench c12_bench_notrackers c12_bench_again c12_bench_notrackers2; do cut -c1-160 $O/$f.json; tail -1 $O/$f.err; done
