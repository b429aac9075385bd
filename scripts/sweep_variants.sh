#!/bin/bash
# import kernel occupancy sweep: variants built by scripts/variants.py
for lib in libipic3d_b200 var_i4 var_i5 var_i6 var_i8 var_b128_8 var_b128_12; do
  echo -n "$lib : "
  IPIC_LIB=$PWD/allscale_ipic3d_b200/lib/$lib.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python scripts/short.py | cut -d'|' -f1-4
done
for lib in libipic3d_b200 var_d4 var_d6 var_d7; do
for st in 2 3 4 6; do
  echo -n "$lib stages=$st : "
  IPIC_STAGES=$st IPIC_LIB=$PWD/allscale_ipic3d_b200/lib/$lib.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python scripts/short.py | cut -d'|' -f1-4
done; done
