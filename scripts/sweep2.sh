#!/bin/bash
for lib in ${LIBS:-libipic3d_b200 var_d4n7 var_d5n7 var_d6n8}; do
for dep in "" "--no-deposit"; do
for items in 1 2; do
for stages in ${STAGES:-2 4 8}; do
  echo -n "$lib items=$items stages=$stages $dep : "
  IPIC_LIB=$PWD/allscale_ipic3d_b200/lib/$lib.so IPIC_ITEMS=$items IPIC_STAGES=$stages python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $dep 2>/dev/null | python scripts/short.py | cut -d'|' -f1-4
done; done; done; done
