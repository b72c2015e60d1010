#!/bin/bash
# GPU call 5: out-of-line chain steps as default, binding evaluations reused at the commit, draw functions out of line
# (A/B); ncu capture of the run launch with sources for the per-function table; launch list of the bench.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_synthetic_systems.py "tests/test_production_parity.py::test_snodin_production_matches_reference_mc" "tests/test_production_parity.py::test_lane_parallel_branches_match_reference_draw_order" -m "gpu and not slow" -x -q) > $O/c5_pytest.log 2>&1
echo "pytest rc=$?" >> $O/c5_pytest.log
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c5_bench.json 2> $O/c5_bench.err
for v in rngout nobindreuse; do
  LDO_B200_LIB=ab/lib_$v.so timeout 300 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c5_bench_$v.json 2> $O/c5_bench_$v.err
done
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c5_bench_again.json 2> $O/c5_bench_again.err
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_exec_staged -o $O/c5_run100 -f python profiles/prof_run.py --moves 100 --replicas 16384 > $O/c5_prof.log 2>&1
timeout 600 ncu --metrics sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__warps_issue_stalled_branch_resolving.avg,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none --profile-from-start off -k regex:k_exec_staged --csv --log-file $O/c5_icc.csv python profiles/prof_run.py --moves 100 --replicas 16384 > $O/c5_prof2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c5_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-regimes > $O/c5_launch_bench.log 2>&1
tail -3 $O/c5_pytest.log; for f in c5_bench c5_bench_rngout c5_bench_nobindreuse c5_bench_again; do cut -c1-160 $O/$f.json; done
