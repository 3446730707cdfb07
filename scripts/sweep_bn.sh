#!/bin/bash
# BatchNorm kernel tuning: unroll x CTAs-per-SM cap, weighted totals of scripts/prof_bn.py
for u in 2 4 8; do
  for cap in 2 4 6 8; do
    echo -n "U=$u cap=$cap  "
    U2_BN_U_STATS=$u U2_BN_U_APPLY=$u U2_BN_U_RED=$u U2_BN_U_BAPPLY=$u U2_BN_CAP_STATS=$cap U2_BN_CAP_APPLY=$cap \
    U2_BN_CAP_RED=$cap U2_BN_CAP_BAPPLY=$cap BN_REPS=5 python scripts/prof_bn.py 2>&1 | tail -1
  done
done
