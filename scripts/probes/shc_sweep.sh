#!/bin/bash
for cfg in "3 32" "1 32" "1 2" "2 2" "1 1" "3 1" "1 4"; do
  set -- $cfg
  echo "== ctas/SM $1 depth $2"
  GS_SHC_CTAS=$1 GS_SHC_DEPTH=$2 AB_REPS=3 AB_STAGES=1 AB_TUNING=0 timeout 200 python scripts/ab_e2e.py 2>&1 | tail -2
done
