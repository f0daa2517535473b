#!/bin/bash
# usage (on the GPU box): tools/ab_cfg3.sh <variant>[:bits] ...  -- config 3 at three operating points per variants/<variant>.so
lib=srslte-emane_b200/libsrslte_b200.so
cp $lib /tmp/default_lib.so
for vb in "$@"; do
  v=${vb%%:*}; bits=0; [[ "$vb" == *:* ]] && bits=${vb##*:}
  echo "== $v bits $bits"; cp variants/$v.so $lib
  for e in 1.5 4.0 6.0; do timeout 60 python tools/cfg3_run.py $e 65536 1 $bits 2>&1 | tail -1 | cut -c1-110; done
done
cp /tmp/default_lib.so $lib
