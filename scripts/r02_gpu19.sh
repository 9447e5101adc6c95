#!/bin/bash
# rows kernels, final instantiation set: whole GPU suite, full-size C4, C4 bench line
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/g19_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/g19_tests.log; tail -n 4 gpurun_out/g19_tests.log
timeout 300 python scripts/full_config.py c4 > gpurun_out/g19_full_c4.json 2> gpurun_out/g19_full_c4.err; echo "full rc=$?"; cut -c1-400 gpurun_out/g19_full_c4.json
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/g19_c4.json 2> gpurun_out/g19_c4.err; echo "bench rc=$?"
python -c "
import json;b=json.load(open('gpurun_out/g19_c4.json'));print(b['value'],b['ms_per_step'],b['config']['variant'][:50],b['roofline']['frac'])"
