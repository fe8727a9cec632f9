#!/usr/bin/env bash
# round-2 GPU call C (2 GPUs: `gpurun --gpus 2`): multi-GPU behind the C ABI -- C++ world-size-2 test, two-process gather test,
# and the bench line at N=2 launched exactly as the driver does
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== comm tests =="; timeout 900 python -m pytest tests/test_gpu_comm.py -m gpu -q 2>&1 | tail -30 | tee $OUT/pytest_comm_r02c.log
echo "== bench N=2 =="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 \
    > $OUT/bench_n2_r02c.json 2> $OUT/bench_n2_r02c.err; tail -5 $OUT/bench_n2_r02c.err; cut -c1-2500 $OUT/bench_n2_r02c.json
