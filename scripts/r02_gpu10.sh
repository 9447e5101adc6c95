#!/bin/bash
# full named sizes, one pass each (scripts/full_config.py): C4 65 536 x 5.4 us, C5 16 384 x 168 ns
mkdir -p gpurun_out
timeout 600 python scripts/full_config.py c4 > gpurun_out/r02_full_c4.json 2> gpurun_out/r02_full_c4.err; echo "c4 rc=$?"; cat gpurun_out/r02_full_c4.json
timeout 900 python scripts/full_config.py c5 > gpurun_out/r02_full_c5.json 2> gpurun_out/r02_full_c5.err; echo "c5 rc=$?"; cat gpurun_out/r02_full_c5.json
