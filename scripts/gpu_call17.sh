#!/bin/bash
# GPU call 17: third batch (hot small functions inlined) - parity on the device and same-box A/B against the second batch.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_synthetic_systems.py -m "gpu and not slow" -x -q) > $O/c17_pytest.log 2>&1
echo "pytest rc=$?" >> $O/c17_pytest.log
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c17_bench.json 2> $O/c17_bench.err
LDO_B200_LIB=ab/lib_merge2.so timeout 300 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c17_bench_merge2.json 2> $O/c17_bench_merge2.err
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c17_bench_again.json 2> $O/c17_bench_again.err
LDO_B200_LIB=ab/lib_merge2.so timeout 300 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c17_bench_merge22.json 2> $O/c17_bench_merge22.err
timeout 600 ncu --metrics sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:k_exec_staged --csv --log-file $O/c17_icc.csv python profiles/prof_run.py --moves 100 --replicas 16384 > $O/c17_prof2.log 2>&1
tail -3 $O/c17_pytest.log; for f in c17_bench c17_bench_merge2 c17_bench_again c17_bench_merge22; do cut -c1-160 $O/$f.json; tail -1 $O/$f.err; done; tail -5 $O/c17_icc.csv | cut -d, -f13-15
