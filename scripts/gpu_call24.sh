#!/bin/bash
# GPU call 24 (the last of the round, 4 GPU-minutes left): the build with the linker movetypes' trackers against its
# twin without them (ab/lib_nolk.so = the device code of the previous build), and the GPU tests the change touches.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
B="--steps 8 --warmup 5 --no-cpu-baseline --no-regimes"
timeout 100 python bench.py $B > $O/c24_bench.json 2> $O/c24_bench.err
LDO_B200_LIB=ab/lib_nolk.so timeout 60 python bench.py $B > $O/c24_bench_nolk.json 2> $O/c24_bench_nolk.err
timeout 60 python bench.py $B > $O/c24_bench_again.json 2> $O/c24_bench_again.err
for f in c24_bench c24_bench_nolk c24_bench_again; do cut -c1-160 $O/$f.json; tail -1 $O/$f.err; done
timeout 100 python -m pytest tests/test_restart_and_outputs.py tests/test_synthetic_systems.py -x -q -m gpu -k "moves_summary or outputs_and_restart or linker" > $O/c24_pytest.log 2>&1
tail -3 $O/c24_pytest.log
