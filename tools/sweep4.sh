#!/bin/bash
for nt in 256 512; do for lm in 3584 5120 7168; do for P in 19 20; do
  echo "== L3_NT=$nt LM=$lm P=$P"
  APGK_L3_NT=$nt APGK_LM=$lm APGK_PREFIX_BITS=$P python tools/prof_run.py 60000000 100000000 25 100 2 2>&1 | tail -1 | sed "s/.*local_max.: [0-9]*} //" | cut -c1-200
done; done; done
