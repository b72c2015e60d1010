#!/bin/bash
# GPU call 25: two instantiations of the kernels (the run kernel of a production launch without the tracker hooks, the
# Tracked<K> one launched while trackers are on) against the one-kernel twin (ab/lib_onekernel.so = device code of the
# build before the linker trackers), then the GPU tests that touch trackers and replay.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
B="--steps 8 --warmup 5 --no-cpu-baseline --no-regimes"
timeout 100 python bench.py $B > $O/c25_bench.json 2> $O/c25_bench.err
LDO_B200_LIB=ab/lib_onekernel.so timeout 60 python bench.py $B > $O/c25_bench_onekernel.json 2> $O/c25_bench_onekernel.err
timeout 60 python bench.py $B > $O/c25_bench_again.json 2> $O/c25_bench_again.err
for f in c25_bench c25_bench_onekernel c25_bench_again; do cut -c1-160 $O/$f.json; tail -1 $O/$f.err; done
timeout 70 python -m pytest tests/test_restart_and_outputs.py tests/test_gpu_parity.py -x -q -m gpu -k "moves_summary or outputs_and_restart or replay_fixture or determinism" > $O/c25_pytest.log 2>&1
tail -3 $O/c25_pytest.log
