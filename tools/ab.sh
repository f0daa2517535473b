#!/bin/bash
# usage (on the GPU box): tools/ab.sh <blocks> <variant> [variant ...]  -- bench_quick per variants/<variant>.so, restoring the default library
lib=srslte-emane_b200/libsrslte_b200.so
cp $lib /tmp/default_lib.so
n=$1; shift
for v in "$@"; do
  echo "== $v"; cp variants/$v.so $lib; bash tools/bench_quick.sh $n
done
cp /tmp/default_lib.so $lib
