#!/bin/bash
# Four-GPU call: the multi-GPU tests (C++ NCCL exchange, CLI with --gpus 2) and the 2- and 4-GPU bench lines of the final build.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/g4_gpus.txt 2>&1
(time timeout 600 python -m pytest tests/test_cli.py tests/test_exchange.py -m gpu -x -q) > $O/g4_pytest.log 2>&1
echo "pytest rc=$?" >> $O/g4_pytest.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 4 > $O/g4_bench2.json 2> $O/g4_bench2.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 8 --warmup 4 > $O/g4_bench4.json 2> $O/g4_bench4.err
tail -3 $O/g4_pytest.log; cut -c1-400 $O/g4_bench2.json; tail -2 $O/g4_bench2.err; cut -c1-400 $O/g4_bench4.json; tail -2 $O/g4_bench4.err
