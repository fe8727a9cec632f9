#!/usr/bin/env bash
# round-2 GPU call O (1 GPU): compute-sanitizer memcheck of the small-batch workload against the final library
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py > $OUT/sanitize_memcheck_r02_final.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/sanitize_memcheck_r02_final.log; tail -4 $OUT/sanitize_memcheck_r02_final.log
