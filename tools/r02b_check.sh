#!/bin/bash
# round-2 (second half) GPU check: async upload test, e2e with / without the asynchronous upload, upload phase laps, bench line
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -5
echo "== e2e sync";  timeout 300 python tools/e2e_jitter.py 2>&1 | tail -3
echo "== e2e async"; timeout 300 python tools/e2e_jitter.py --async 2>&1 | tail -3
echo "== upload laps (sync)"; FSPT_TIMING=1 timeout 300 python tools/upload_time.py 2>&1 | tail -40
} > gpurun_out/r02b_check.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench_c2.json 2> gpurun_out/r02b_bench_c2.err
tail -3 gpurun_out/r02b_check.log; cat gpurun_out/r02b_bench_c2.json | head -c 1500
