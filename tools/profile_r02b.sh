#!/bin/bash
# Round 2, second half (host-side work: asynchronous / pipelined scene upload; kernels unchanged since tools/profile_r02.sh):
# full GPU suite, smoke, one bench line per BASELINE config (value, e2e, roofline, parity), the reference arm, upload laps.
# Everything lands in gpurun_out/r02b_*; `python profiles/refresh.py r02b` copies the lines into profiles/.
# usage: tools/profile_r02b.sh [configs, default "2 1 3 4 5"]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CONFIGS="${1:-2 1 3 4 5}"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r02b_gputest.log; cat gpurun_out/r02b_gputest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee gpurun_out/r02b_smoke.log
for c in $CONFIGS; do
  steps=10; [ "$c" = "5" ] && steps=3; [ "$c" = "4" ] && steps=5; [ "$c" = "2" ] && steps=20
  extra=""; [ "$c" != "2" ] && extra="--no-cpu-baseline"
  timeout 900 python bench.py --config $c --steps $steps --warmup 3 $extra > gpurun_out/r02b_bench_c$c.json 2> gpurun_out/r02b_bench_c$c.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02b_bench_c$c.json"))
    print("config $c: value %.1f (%.2f ms)  e2e %.1f (%.2f ms)  roofline %s %.3f  parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["bound"], d["roofline"]["frac"], d.get("parity")))
except Exception as e:
    print("config $c: no line (%s)" % e)
PY
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02b_bench_reference.json 2> gpurun_out/r02b_bench_reference.err
head -c 400 gpurun_out/r02b_bench_reference.json; echo
{
echo "== e2e sync";  timeout 300 python tools/e2e_jitter.py 2>&1 | tail -3
echo "== e2e async"; timeout 300 python tools/e2e_jitter.py --async 2>&1 | tail -3
echo "== upload laps (async, config 2 scene)"; FSPT_TIMING=1 timeout 300 python tools/upload_time.py --async 2>&1 | tail -22 | head -18
echo "== upload laps (sync, config 3 scene: 1 M triangles)"; FSPT_TIMING=1 timeout 300 python tools/upload_time.py --soup 2>&1 | tail -13 | head -9
} > gpurun_out/r02b_upload.log 2>&1
tail -12 gpurun_out/r02b_upload.log
# ncu evidence of the final kernels (80-byte path records), config 2: launch list + one --set full capture per kernel
if [ "${NCU:-1}" = "1" ]; then
  NCUCMD="ncu --clock-control none"
  BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-verify"
  timeout 900 $NCUCMD --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02b_launches.csv $BENCH > gpurun_out/r02b_launches.log 2>&1
  timeout 1200 $NCUCMD --set full --import-source on -k regex:k_trace --launch-skip 15 --launch-count 3 -f -o gpurun_out/r02b_k_trace $BENCH > gpurun_out/r02b_ncu_trace.log 2>&1
  timeout 1200 $NCUCMD --set full --import-source on -k regex:k_shade --launch-skip 15 --launch-count 2 -f -o gpurun_out/r02b_k_shade $BENCH > gpurun_out/r02b_ncu_shade.log 2>&1
  ls -la gpurun_out/ | grep -E "r02b_(k_|launches)"
fi
