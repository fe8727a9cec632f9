#!/usr/bin/env bash
# round-2 GPU call I (8 GPUs): the bench line at N=8 with the peer transport (the default), the driver's flags; then NCCL, C2 only
set -u
OUT=gpurun_out
mkdir -p $OUT
show() { grep '^{' $1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','ms_per_step','step_ms','transport','peer_status','parity','kernel_ms')}); print({k:(d.get(k) or {}).get('ms_per_step', (d.get(k) or {}).get('ms_per_batch')) for k in ('c3','c5')})"; }
echo "== bench N=8, peer transport (driver's flags) =="
PHANT_BENCH_TRANSPORT=peer timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 \
    > $OUT/bench_n8_peer_r02i.json 2> $OUT/bench_n8_peer_r02i.err; tail -3 $OUT/bench_n8_peer_r02i.err; show $OUT/bench_n8_peer_r02i.json
echo "== bench N=8, NCCL, C2 only =="
PHANT_BENCH_TRANSPORT=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 5 --skip-extras \
    > $OUT/bench_n8_nccl_r02i.json 2> $OUT/bench_n8_nccl_r02i.err; tail -3 $OUT/bench_n8_nccl_r02i.err; show $OUT/bench_n8_nccl_r02i.json
echo "== bench N=1 same box, C2 only =="
timeout 200 python bench.py --steps 20 --warmup 5 --skip-extras --no-cpu > $OUT/bench_n1_r02i.json 2> $OUT/bench_n1_r02i.err; show $OUT/bench_n1_r02i.json
