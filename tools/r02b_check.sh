#!/bin/bash
# round-2 (second half) GPU check: full GPU suite, e2e with / without the asynchronous upload, upload phase laps, bench lines
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -25
echo "== e2e sync";  timeout 300 python tools/e2e_jitter.py 2>&1 | tail -3
echo "== e2e async"; timeout 300 python tools/e2e_jitter.py --async 2>&1 | tail -3
echo "== upload laps (sync)"; FSPT_TIMING=1 timeout 300 python tools/upload_time.py 2>&1 | tail -42
} > gpurun_out/r02b_check.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench_c2.json 2> gpurun_out/r02b_bench_c2.err
timeout 900 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_c3.json 2> gpurun_out/r02b_bench_c3.err
tail -3 gpurun_out/r02b_check.log; head -c 600 gpurun_out/r02b_bench_c2.json
