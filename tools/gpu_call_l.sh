#!/usr/bin/env bash
# round-2 GPU call L (2 GPUs): final state -- the whole -m gpu suite (incl. the two-GPU tests), smoke, the full bench line at N=2 and N=1
set -u
OUT=gpurun_out
mkdir -p $OUT
show() { grep '^{' $1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','ms_per_step','transport','parity','kernel_ms')}); print({k:(d.get(k) or {}).get('ms_per_step', (d.get(k) or {}).get('ms_per_batch')) for k in ('c3','c5')}, (d.get('c5') or {}).get('witness_build_s_on_device'), (d.get('c4_sparse') or {}).get('ms_per_update'))"; }
echo "== gpu tests =="; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $OUT/pytest_gpu_r02l.log
echo "== smoke ==";     timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke_r02l.log
echo "== bench N=2 =="
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 \
    > $OUT/bench_n2_r02l.json 2> $OUT/bench_n2_r02l.err; tail -3 $OUT/bench_n2_r02l.err; show $OUT/bench_n2_r02l.json
echo "== bench N=1 =="
timeout 700 python bench.py > $OUT/bench_n1_r02l.json 2> $OUT/bench_n1_r02l.err; tail -3 $OUT/bench_n1_r02l.err; show $OUT/bench_n1_r02l.json
