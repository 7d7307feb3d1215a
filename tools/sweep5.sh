#!/bin/bash
for cfg in "17 9" "18 9" "18 10" "19 9" "19 10" "19 11"; do set -- $cfg
  echo "== P=$1 D0=$2"
  APGK_PREFIX_BITS=$1 APGK_D0=$2 python tools/prof_run.py 60000000 100000000 25 100 2 2>&1 | tail -1 | sed "s/.*local_max.: [0-9]*} //" | cut -c1-200
done
