#!/bin/bash
# attention microbench under arbitrary env settings: each argument is a "VAR=val,VAR2=val" list
cd "$(dirname "$0")/.."
for cfg in "$@"; do
  echo "== $cfg"
  env ${cfg//,/ } timeout 300 python tools/microbench.py --only attn 2>&1 | grep "attn"
done
