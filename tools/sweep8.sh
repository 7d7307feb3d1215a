#!/bin/bash
# compare L2-hint / pipelining build variants of the scatter kernels (libapgk_<tag>.so built beside libapgk.so)
cp allpathslg_b200/libapgk.so /tmp/libapgk_default.so
for tag in default "$@"; do
  if [ "$tag" != "default" ]; then cp allpathslg_b200/libapgk_$tag.so allpathslg_b200/libapgk.so; else cp /tmp/libapgk_default.so allpathslg_b200/libapgk.so; fi
  echo "== $tag"
  timeout -k 5 90 python tools/prof_run.py 60000000 100000000 25 100 2 2>&1 | tail -1 | sed 's/.*n_rounds.: [0-9]*} //' | cut -c1-230
done
cp /tmp/libapgk_default.so allpathslg_b200/libapgk.so
