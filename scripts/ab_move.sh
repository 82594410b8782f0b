#!/bin/bash
# A/B of library variants on the MH fold: scripts/ab_move.sh name1 name2 ...   ("main" = the in-tree library)
for v in "$@"; do
  if [ "$v" = main ]; then unset WSB200_LIB; else export WSB200_LIB=$PWD/variants/$v.so; fi
  echo "== $v"
  python scripts/prof_move.py 2000 10000000 2>&1 | tail -2
  python benchmarks/run_configs.py c3 hier 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['config'][:60], d['seconds'])"
done
