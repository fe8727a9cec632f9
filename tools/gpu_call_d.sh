#!/usr/bin/env bash
# round-2 GPU call D (8 GPUs: `gpurun --gpus 8`): the bench line at N=8 exactly as the driver launches it, then a long C2-only
# run (100 steps) for the per-step min / median / max
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== bench N=8 (driver's flags) =="
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 \
    > $OUT/bench_n8_r02d.json 2> $OUT/bench_n8_r02d.err; tail -5 $OUT/bench_n8_r02d.err; cut -c1-2500 $OUT/bench_n8_r02d.json
echo "== bench N=8, 100 steps, C2 only =="
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 100 --warmup 5 --skip-extras \
    > $OUT/bench_n8_100_r02d.json 2> $OUT/bench_n8_100_r02d.err; tail -3 $OUT/bench_n8_100_r02d.err; cut -c1-1500 $OUT/bench_n8_100_r02d.json
echo "== bench N=1 on the same box, C2 only =="
timeout 300 python bench.py --steps 20 --warmup 5 --skip-extras --no-cpu > $OUT/bench_n1_r02d.json 2> $OUT/bench_n1_r02d.err; cut -c1-600 $OUT/bench_n1_r02d.json
