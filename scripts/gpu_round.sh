#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list.  Usage: gpurun -- 'bash scripts/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $OUT/clocks_$TAG.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -5 $OUT/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?"; cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
echo "ref exit $?"; cat $OUT/bench_ref_$TAG.json
kill $SMI
if [ "$2" != "noncu" ]; then
ALG_BENCH_FAST=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 1 --warmup 0 > $OUT/ncu_bench_$TAG.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/ncu_bench_$TAG.log; wc -l $OUT/launches_$TAG.csv
fi
