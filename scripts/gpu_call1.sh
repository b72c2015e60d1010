#!/bin/bash
# GPU call 1 of round 2: parity first, then the stationary bench, A/B of this round's changes, exploration.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/c1_gpu.txt 2>&1
(time timeout 1500 python -m pytest tests -m "gpu and not slow" -x -q -s) > $O/c1_pytest.log 2>&1
echo "pytest rc=$?" >> $O/c1_pytest.log
timeout 400 python bench.py --steps 6 --warmup 5 --no-cpu-baseline > $O/c1_bench_s6.json 2> $O/c1_bench_s6.err
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-regimes > $O/c1_bench_s40.json 2> $O/c1_bench_s40.err
LDO_B200_LIB=ab/lib_noweightpass.so timeout 400 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > $O/c1_bench_noweightpass.json 2> $O/c1_bench_noweightpass.err
LDO_ORDER_MODE=grouped timeout 400 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-regimes > $O/c1_bench_grouped.json 2> $O/c1_bench_grouped.err
timeout 400 python bench.py --workload ct_four --steps 6 --warmup 3 --no-cpu-baseline > $O/c1_bench_ct_four.json 2> $O/c1_bench_ct_four.err
timeout 400 python bench.py --workload ptmwus --steps 6 --warmup 3 --no-cpu-baseline > $O/c1_bench_ptmwus.json 2> $O/c1_bench_ptmwus.err
timeout 300 python profiles/relax_four.py 2048 300000 20000 > $O/c1_relax_four.txt 2>&1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/c1_bench_ref.json 2> $O/c1_bench_ref.err
tail -3 $O/c1_pytest.log; cat $O/c1_bench_s6.json | cut -c1-600; cat $O/c1_bench_s40.json | cut -c1-300
