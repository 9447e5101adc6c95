#!/bin/bash
# runs every golden dc/tran case with a per-case timeout
cd "$(dirname "$0")/.."
for c in "v_divider_sweep dc" "proj2 dc" "i_divider_sweep dc" "rc tran" "rl tran" "rc_sine tran" "rc_exp tran" "rc_pulse tran" "proj3 tran" "rc_i_sine tran"; do
  timeout 40 python scripts/dbg_cases.py $c 2>&1 | tail -5 || echo "TIMEOUT/FAIL: $c"
done
