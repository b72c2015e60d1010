#!/bin/bash
# Two-GPU call: the NCCL exchange inside the C++ host (CLI with --gpus 2 against --gpus 1) and the 2-GPU bench line.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/g2_gpus.txt 2>&1
(time timeout 900 python -m pytest tests/test_cli.py tests/test_us_oracle.py tests/test_umbrella_sampling.py tests/test_exchange.py -m gpu -x -q) > $O/g2_pytest.log 2>&1
echo "pytest rc=$?" >> $O/g2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 5 > $O/g2_bench.json 2> $O/g2_bench.err
./profiles/smem_bw > $O/g2_smem.json 2>&1
tail -3 $O/g2_pytest.log; cat $O/g2_smem.json; cut -c1-500 $O/g2_bench.json; tail -5 $O/g2_bench.err
