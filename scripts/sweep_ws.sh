#!/bin/bash
# warp-specialised mover: stages sweep of the product build, and the four-warp kernel (IPIC_WS=0) beside it
for st in 4 6 8; do
  echo -n "k_move_ws stages=$st : "
  IPIC_STAGES=$st python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python scripts/short.py | cut -d'|' -f1-4
done
echo -n "k_move (IPIC_WS=0) : "; IPIC_WS=0 IPIC_STAGES=4 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python scripts/short.py | cut -d'|' -f1-4
