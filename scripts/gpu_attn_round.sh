#!/bin/bash
# attention tuning visit: variants + ncu full capture.
# Usage: gpurun -- 'bash scripts/gpu_attn_round.sh tag "0 2 4 8" "0 1" [noncu]'   (poly list, split list)
TAG=${1:-a01}
OUT=gpurun_out
mkdir -p $OUT
for sp in ${3:-0 1}; do
for v in ${2:-0 4}; do
  ALG_ATTN_SPLIT=$sp ALG_ATTN_POLY=$v timeout 300 python scripts/attn_bench.py 10 40 2>&1 | tail -1 | sed "s/^/split=$sp /" | tee -a $OUT/attn_$TAG.log
done
done
if [ "$4" != "noncu" ]; then
for sp in ${3:-0 1}; do
ALG_ATTN_SPLIT=$sp HEADS=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 1 -c 1 -f -o $OUT/attn_${TAG}_split$sp python scripts/prof_kernels.py attn > $OUT/ncu_attn_${TAG}_split$sp.log 2>&1
tail -2 $OUT/ncu_attn_${TAG}_split$sp.log
done
fi
