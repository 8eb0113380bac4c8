#!/bin/bash
# A/B of build variants on one box: tools/ab.sh lib1.so lib2.so ...   (paths relative to operator-builder_b200/; "-" = the default library)
for rep in 1 2; do
for v in "$@"; do
  if [ "$v" = "-" ]; then L=""; else L="$PWD/operator-builder_b200/$v"; fi
  echo -n "$v: "; OBM_LIB=$L python tools/profile_run.py --docs 262144 --iters 6 2>&1 | sort -t' ' -k3 -n | grep iter | awk '{print $3}' | sort -n | head -1
done; done
