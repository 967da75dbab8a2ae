#!/bin/bash
# Round-2 evidence pass (run on the GPU box through gpurun; everything lands in gpurun_out/, the summaries are then
# copied into profiles/ by profiles/refresh.py on the CPU box):
#   1. one bench line per BASELINE config (value, e2e, roofline, parity)          -> r02_bench_c<N>.json
#   2. the ncu launch list of the default bench command (per-launch durations)    -> r02_launches.csv
#   3. one `ncu --set full` capture of the traversal kernel (primary launch + first two bounce launches) and of the
#      shading kernel, config 2                                                   -> r02_k_trace.ncu-rep, r02_k_shade.ncu-rep
#   4. the same traversal capture on config 5 (10 M triangles, 4K: HBM-bound)     -> r02_k_trace_c5.ncu-rep
# usage: tools/profile_r02.sh [configs, default "2 1 3 4 5"]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CONFIGS="${1:-2 1 3 4 5}"
for c in $CONFIGS; do
  steps=10; [ "$c" = "5" ] && steps=3; [ "$c" = "4" ] && steps=5
  extra=""; [ "$c" != "2" ] && extra="--no-cpu-baseline"
  timeout 900 python bench.py --config $c --steps $steps --warmup 3 $extra > gpurun_out/r02_bench_c$c.json 2> gpurun_out/r02_bench_c$c.err
  tail -c 600 gpurun_out/r02_bench_c$c.json; echo
done
NCU="ncu --clock-control none"
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-verify"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_launches.csv $BENCH > gpurun_out/r02_launches.log 2>&1
timeout 1200 $NCU --set full --import-source on -k regex:k_trace --launch-skip 15 --launch-count 3 -f -o gpurun_out/r02_k_trace $BENCH > gpurun_out/r02_ncu_trace.log 2>&1
timeout 1200 $NCU --set full --import-source on -k regex:k_shade --launch-skip 15 --launch-count 2 -f -o gpurun_out/r02_k_shade $BENCH > gpurun_out/r02_ncu_shade.log 2>&1
if echo "$CONFIGS" | grep -q 5; then
  timeout 1500 $NCU --set full --import-source on -k regex:k_trace --launch-skip 5 --launch-count 2 -f -o gpurun_out/r02_k_trace_c5 \
    python bench.py --config 5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-verify > gpurun_out/r02_ncu_trace_c5.log 2>&1
fi
ls -la gpurun_out/ | grep r02_
