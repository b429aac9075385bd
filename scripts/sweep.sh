#!/bin/bash
# tuning sweep of the mover's tile shape: items x stages, with and without deposit
for dep in "" "--no-deposit"; do
for items in 1 2; do
for stages in 1 2 3 4 6; do
  echo -n "items=$items stages=$stages $dep : "
  IPIC_ITEMS=$items IPIC_STAGES=$stages python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $dep 2>/dev/null | python scripts/short.py | cut -d'|' -f1-4
done; done; done
