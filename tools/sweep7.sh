#!/bin/bash
cp allpathslg_b200/libapgk.so /tmp/libapgk_default.so
for lib in libapgk.so libapgk_nt512.so; do
  if [ "$lib" != "libapgk.so" ]; then cp allpathslg_b200/$lib allpathslg_b200/libapgk.so; fi
  for ct in 16 128; do
    echo "== $lib CT=$ct"
    APGK_CT=$ct python tools/prof_run.py 60000000 100000000 25 100 2 2>&1 | tail -1 | sed 's/.*n_rounds.: [0-9]*} //' | cut -c1-210
  done
done
cp /tmp/libapgk_default.so allpathslg_b200/libapgk.so
