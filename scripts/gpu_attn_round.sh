#!/bin/bash
# attention tuning visit: variants + ncu full capture.  Usage: gpurun -- 'bash scripts/gpu_attn_round.sh tag "0 2 4 8"'
TAG=${1:-a01}
OUT=gpurun_out
mkdir -p $OUT
for v in ${2:-0 4}; do
  ALG_ATTN_POLY=$v timeout 300 python scripts/attn_bench.py 10 40 2>&1 | tail -1 | tee -a $OUT/attn_$TAG.log
done
if [ "$3" != "noncu" ]; then
HEADS=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 1 -c 1 -o $OUT/attn_$TAG python scripts/prof_kernels.py attn > $OUT/ncu_attn_$TAG.log 2>&1
tail -2 $OUT/ncu_attn_$TAG.log
fi
