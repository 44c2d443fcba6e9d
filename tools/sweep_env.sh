#!/bin/bash
# usage: tools/sweep_env.sh VAR v1 v2 ... -- [bench args]   -> one line per value: value, device pairs/s, e2e pairs/s
var=$1; shift
vals=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do vals+=("$1"); shift; done
shift
for v in "${vals[@]}"; do
  line=$(env "$var=$v" timeout 150 python bench.py "$@" 2>/dev/null | tail -1)
  echo "$line" | V="$var=$v" python -c 'import sys,json,os; d=json.loads(sys.stdin.read()); print(os.environ["V"], round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"])'
done
