#!/bin/bash
# GPU call 3: full GPU test-suite (with the new rows), bench line with the reference beside it, build variants,
# shared-memory peak, sanitizer runs, the slow equilibrium test.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
(time timeout 1500 python -m pytest tests -m "gpu and not slow" -x -q) > $O/c3_pytest.log 2>&1
echo "pytest rc=$?" >> $O/c3_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/c3_bench.json 2> $O/c3_bench.err
for v in ptxasO2 mb24 mb20; do
  LDO_B200_LIB=ab/lib_$v.so timeout 300 python bench.py --steps 6 --warmup 4 --no-cpu-baseline > $O/c3_bench_$v.json 2> $O/c3_bench_$v.err
done
./profiles/smem_bw > $O/c3_smem.json 2>&1
(time timeout 700 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/sanitize_run.py) > $O/c3_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/c3_memcheck.log
(time timeout 700 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitize_run.py) > $O/c3_racecheck.log 2>&1
echo "racecheck rc=$?" >> $O/c3_racecheck.log
(time timeout 900 python -m pytest tests/test_production_parity.py -m "gpu and slow" -x -q -s) > $O/c3_slow.log 2>&1
echo "slow rc=$?" >> $O/c3_slow.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c3_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-regimes > $O/c3_launch_bench.log 2>&1
tail -3 $O/c3_pytest.log; cut -c1-400 $O/c3_bench.json; tail -2 $O/c3_memcheck.log; tail -2 $O/c3_racecheck.log; tail -4 $O/c3_slow.log
