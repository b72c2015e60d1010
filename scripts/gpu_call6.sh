#!/bin/bash
# GPU call 6: compact build (draw functions and small helpers out of line, snapshot revert of rejected moves): full GPU
# suite, A/B against its inline-helper and reference-revert twins and the previous build (cross-box anchor), ncu capture.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
(time timeout 1500 python -m pytest tests -m "gpu and not slow" -x -q) > $O/c6_pytest.log 2>&1
echo "pytest rc=$?" >> $O/c6_pytest.log
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c6_bench.json 2> $O/c6_bench.err
for v in inlinehelpers nofastrevert stepoutline; do
  LDO_B200_LIB=ab/lib_$v.so timeout 300 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c6_bench_$v.json 2> $O/c6_bench_$v.err
done
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c6_bench_again.json 2> $O/c6_bench_again.err
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_exec_staged -o $O/c6_run100 -f python profiles/prof_run.py --moves 100 --replicas 16384 > $O/c6_prof.log 2>&1
timeout 600 ncu --metrics sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__warps_issue_stalled_branch_resolving.avg,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none --profile-from-start off -k regex:k_exec_staged --csv --log-file $O/c6_icc.csv python profiles/prof_run.py --moves 100 --replicas 16384 > $O/c6_prof2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c6_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-regimes > $O/c6_launch_bench.log 2>&1
tail -3 $O/c6_pytest.log; for f in c6_bench c6_bench_inlinehelpers c6_bench_nofastrevert c6_bench_stepoutline c6_bench_again; do cut -c1-160 $O/$f.json; done
