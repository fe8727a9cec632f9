#!/usr/bin/env bash
# round-2 GPU call M (2 GPUs): the full bench line at N=2 (peer transport) and N=1 after the generator's stream-ordering fix
set -u
OUT=gpurun_out
mkdir -p $OUT
show() { grep '^{' $1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','ms_per_step','transport','peer_status','parity','kernel_ms')}); print({k:(d.get(k) or {}).get('ms_per_step', (d.get(k) or {}).get('ms_per_batch')) for k in ('c3','c5')}, (d.get('c5') or {}).get('witness_build_s_on_device'), (d.get('c5') or {}).get('parity'), (d.get('c4_sparse') or {}).get('ms_per_update'))"; }
echo "== bench N=2 =="
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 \
    > $OUT/bench_n2_r02m.json 2> $OUT/bench_n2_r02m.err; tail -3 $OUT/bench_n2_r02m.err | cut -c1-300; show $OUT/bench_n2_r02m.json
echo "== bench N=1 =="
timeout 700 python bench.py > $OUT/bench_n1_r02m.json 2> $OUT/bench_n1_r02m.err; tail -3 $OUT/bench_n1_r02m.err | cut -c1-300; show $OUT/bench_n1_r02m.json
echo "== C5 generator test =="; timeout 300 python -m pytest tests/test_gpu_verify.py -m gpu -q -k "device_built" 2>&1 | tail -2
