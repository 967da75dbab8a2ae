#!/bin/bash
# round-2 (second half) GPU check: GPU suite, e2e with / without the asynchronous upload, upload phase laps, bench lines
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -x -q -m gpu -k "not ten_million" 2>&1 | tail -4
echo "== e2e sync";  timeout 300 python tools/e2e_jitter.py 2>&1 | tail -3
echo "== e2e async"; timeout 300 python tools/e2e_jitter.py --async 2>&1 | tail -3
echo "== e2e sync, pinned atlas + env"; timeout 300 python tools/e2e_jitter.py --pinned 2>&1 | tail -3
echo "== e2e async, pinned atlas + env"; timeout 300 python tools/e2e_jitter.py --async --pinned 2>&1 | tail -3
echo "== upload laps (async, bunny)"; FSPT_TIMING=1 timeout 300 python tools/upload_time.py --async 2>&1 | tail -22 | head -18
echo "== upload laps (async, pinned, bunny)"; FSPT_TIMING=1 timeout 300 python tools/upload_time.py --async --pinned 2>&1 | tail -22 | head -18
} > gpurun_out/r02b_check.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench_c2.json 2> gpurun_out/r02b_bench_c2.err
tail -3 gpurun_out/r02b_check.log; head -c 300 gpurun_out/r02b_bench_c2.json
