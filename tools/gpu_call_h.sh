#!/usr/bin/env bash
# round-2 GPU call H (2 GPUs): the peer transport -- tests, then the bench line with it and with NCCL
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== comm tests =="; timeout 600 python -m pytest tests/test_gpu_comm.py -m gpu -q 2>&1 | tail -30 | tee $OUT/pytest_comm_r02h.log
echo "== bench N=2, peer transport =="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --skip-extras \
    > $OUT/bench_n2_peer_r02h.json 2> $OUT/bench_n2_peer_r02h.err; tail -3 $OUT/bench_n2_peer_r02h.err; grep '^{' $OUT/bench_n2_peer_r02h.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','step_ms','transport','peer_status','parity','kernel_ms')})"
echo "== bench N=2, NCCL =="
PHANT_BENCH_TRANSPORT=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 --skip-extras \
    > $OUT/bench_n2_nccl_r02h.json 2> $OUT/bench_n2_nccl_r02h.err; tail -3 $OUT/bench_n2_nccl_r02h.err; grep '^{' $OUT/bench_n2_nccl_r02h.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','step_ms','transport','parity','kernel_ms')})"
