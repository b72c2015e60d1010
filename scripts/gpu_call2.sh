#!/bin/bash
# GPU call 2: lane-parallel FourBody potential - parity, A/B against the serial evaluation, ncu of the run launch.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_synthetic_systems.py "tests/test_production_parity.py::test_snodin_production_matches_reference_mc" -m "gpu and not slow" -x -q) > $O/c2_pytest.log 2>&1
echo "pytest rc=$?" >> $O/c2_pytest.log
timeout 400 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > $O/c2_bench.json 2> $O/c2_bench.err
LDO_B200_LIB=ab/lib_serialpot.so timeout 400 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > $O/c2_bench_serialpot.json 2> $O/c2_bench_serialpot.err
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_exec_staged -o $O/c2_run100 -f python profiles/prof_run.py --moves 100 --replicas 16384 > $O/c2_prof.log 2>&1
timeout 600 ncu --metrics sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__warps_issue_stalled_branch_resolving.avg,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none --profile-from-start off -k regex:k_exec_staged --csv --log-file $O/c2_icc.csv python profiles/prof_run.py --moves 100 --replicas 16384 > $O/c2_prof2.log 2>&1
tail -3 $O/c2_pytest.log; cut -c1-300 $O/c2_bench.json; cut -c1-300 $O/c2_bench_serialpot.json
