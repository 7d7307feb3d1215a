#!/bin/bash
# usage: tools/sweep.sh  -- runs prof_run over (lib variant, CT, P) and prints stage times
for lib in libapgk.so libapgk_nt1024.so; do
  cp allpathslg_b200/libapgk.so /tmp/libapgk_default.so
  if [ "$lib" != "libapgk.so" ]; then cp allpathslg_b200/$lib allpathslg_b200/libapgk.so; fi
  for P in 21 22; do for ct in 1 8 32 128; do
    echo "== $lib P=$P CT=$ct"
    APGK_PREFIX_BITS=$P APGK_CT=$ct python tools/prof_run.py 30000000 50000000 25 100 2 2>&1 | tail -1 | sed 's/.*local_max.: [0-9]*} //'
  done; done
  cp /tmp/libapgk_default.so allpathslg_b200/libapgk.so
done
