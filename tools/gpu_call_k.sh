#!/usr/bin/env bash
# round-2 GPU call K (2 GPUs): bench.py after the transport refactor (N=2 default = peer, N=1), sparse-trie phase trace
set -u
OUT=gpurun_out
mkdir -p $OUT
show() { grep '^{' $1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','ms_per_step','transport','peer_status','parity','kernel_ms')})"; }
echo "== bench N=2 (default transport) =="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --skip-extras \
    > $OUT/bench_n2_r02k.json 2> $OUT/bench_n2_r02k.err; tail -3 $OUT/bench_n2_r02k.err; show $OUT/bench_n2_r02k.json
echo "== bench N=1 =="
timeout 300 python bench.py --steps 20 --warmup 5 --skip-extras --no-cpu > $OUT/bench_n1_r02k.json 2> $OUT/bench_n1_r02k.err; tail -3 $OUT/bench_n1_r02k.err; show $OUT/bench_n1_r02k.json
echo "== sparse trie phase trace (16.7M keys, 100k upserts) =="
PHANT_GPU_TRACE=1 timeout 300 python tools/strie_bench.py --keys 16777216 --dirty 100000 --steps 2 2> $OUT/strie_trace_r02k.log | tail -1; tail -24 $OUT/strie_trace_r02k.log
echo "== comm tests =="; timeout 300 python -m pytest tests/test_gpu_comm.py -m gpu -q 2>&1 | tail -4 | tee $OUT/pytest_comm_r02k.log
