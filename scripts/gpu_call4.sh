#!/bin/bash
# GPU call 4: slot / placement caches and inline chain steps - parity on the device and A/B.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_synthetic_systems.py "tests/test_production_parity.py::test_snodin_production_matches_reference_mc" "tests/test_production_parity.py::test_lane_parallel_branches_match_reference_draw_order" -m "gpu and not slow" -x -q) > $O/c4_pytest.log 2>&1
echo "pytest rc=$?" >> $O/c4_pytest.log
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c4_bench.json 2> $O/c4_bench.err
for v in noslotcache stepoutline; do
  LDO_B200_LIB=ab/lib_$v.so timeout 300 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c4_bench_$v.json 2> $O/c4_bench_$v.err
done
timeout 400 python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-regimes > $O/c4_bench_again.json 2> $O/c4_bench_again.err
tail -3 $O/c4_pytest.log; for f in c4_bench c4_bench_noslotcache c4_bench_stepoutline c4_bench_again; do cut -c1-160 $O/$f.json; done
