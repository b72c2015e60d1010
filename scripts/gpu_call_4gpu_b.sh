#!/bin/bash
# Four-GPU call (final build): the 2- and 4-GPU bench lines.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 4 --no-regimes > $O/g4b_bench2.json 2> $O/g4b_bench2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 8 --warmup 4 --no-regimes > $O/g4b_bench4.json 2> $O/g4b_bench4.err
timeout 300 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-regimes > $O/g4b_bench1.json 2> $O/g4b_bench1.err
cut -c1-300 $O/g4b_bench2.json; tail -2 $O/g4b_bench2.err; cut -c1-300 $O/g4b_bench4.json; tail -2 $O/g4b_bench4.err; cut -c1-200 $O/g4b_bench1.json
