#!/bin/bash
# GPU call 19 (final build of the round): whole GPU suite, smoke(), the default bench line with cpu_baseline, the reference arm, A/B of the
# last two changes on the same box, ncu capture + instruction-cache counters + launch list of the final build.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
(time timeout 1800 python -m pytest tests -m "gpu and not slow" -x -q) > $O/c19_pytest.log 2>&1
echo "pytest rc=$?" >> $O/c19_pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > $O/c19_smoke.log 2>&1
echo "smoke rc=$?" >> $O/c19_smoke.log
timeout 900 python bench.py > $O/c19_bench.json 2> $O/c19_bench.err
timeout 900 python bench.py --impl reference > $O/c19_bench_ref.json 2> $O/c19_bench_ref.err
for v in splitsingle; do
  LDO_B200_LIB=ab/lib_$v.so timeout 300 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c19_bench_$v.json 2> $O/c19_bench_$v.err
done
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c19_bench_again.json 2> $O/c19_bench_again.err
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_exec_staged -o $O/c19_run100 -f python profiles/prof_run.py --moves 100 --replicas 16384 > $O/c19_prof.log 2>&1
timeout 600 ncu --metrics sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__warps_issue_stalled_branch_resolving.avg,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none --profile-from-start off -k regex:k_exec_staged --csv --log-file $O/c19_icc.csv python profiles/prof_run.py --moves 100 --replicas 16384 > $O/c19_prof2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c19_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-regimes > $O/c19_launch_bench.log 2>&1
tail -3 $O/c19_pytest.log; tail -1 $O/c19_smoke.log; for f in c19_bench c19_bench_ref c19_bench_splitsingle c19_bench_again; do cut -c1-160 $O/$f.json; tail -1 $O/$f.err; done
