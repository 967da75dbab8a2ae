#!/bin/bash
# 8-GPU box: weak / strong scaling lines of bench.py, config 2 (outputs under gpurun_out/r02b_scale_*.json)
cd "$(dirname "$0")/.."
run() {  # n tag args...
  n=$1; tag=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus $n "$@" > gpurun_out/r02b_scale_$tag.json 2> gpurun_out/r02b_scale_$tag.err
  echo "== $tag rc=$?"; grep '^{' gpurun_out/r02b_scale_$tag.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'scaling', 'parity')}, 'e2e', round(d['e2e']['value'], 1), round(d['e2e']['ms_per_step'], 2), d['config']['parallelism'])"
  tail -3 gpurun_out/r02b_scale_$tag.err | cut -c1-300
}
run 8 n8_weak --steps 10 --warmup 3
run 8 n8_strong --steps 10 --warmup 3 --scaling strong
run 4 n4_weak --steps 10 --warmup 3
