#!/usr/bin/env bash
# round-2 GPU call G (1 GPU, short): sparse resident trie with the leaf-reference cache
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== tests =="; timeout 900 python -m pytest tests/test_gpu_trie.py tests/test_gpu_host_py.py tests/test_gpu_host_cpp.py -m gpu -q 2>&1 | tail -6 | tee $OUT/pytest_r02g.log
echo "== sparse trie =="; timeout 300 python tools/strie_bench.py --keys 4000000 --dirty 100000 --steps 5 | tee $OUT/strie_r02g.json
timeout 300 python tools/strie_bench.py --keys 16777216 --dirty 100000 --steps 5 | tee -a $OUT/strie_r02g.json
timeout 300 python tools/builders_bench.py 2>&1 | tail -12 | cut -c1-220 | tee $OUT/builders_r02g.log
