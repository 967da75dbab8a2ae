#!/bin/bash
# usage: tools/ab_run.sh "<scenes>" <spp> variant[:ENV=VAL] ...   ("." = the default build); prints one line per run
scenes="$1"; spp="$2"; shift 2
for spec in "$@"; do
  v="${spec%%:*}"; envs=""
  [ "$spec" != "$v" ] && envs="${spec#*:}"
  lib=""; [ "$v" != "." ] && lib="fspt_b200/lib/variants/$v.so"
  for s in $scenes; do
    echo -n "$spec $s | "
    env ${envs//,/ } FSPT_LIB=$lib timeout 300 python tools/quick.py --scene $s --reps 2 --spp $spp 2>&1 | tail -1
  done
done
