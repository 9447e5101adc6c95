#!/bin/bash
mkdir -p gpurun_out
for fc in 800 1200 2000; do timeout 300 python scripts/dbg_c5_op.py 4096 fill_cap=$fc 2>&1 | tail -1 | cut -c1-420; done
timeout 300 python scripts/dbg_c5_op.py 4096 sparse_warp=0 2>&1 | tail -1 | cut -c1-420
