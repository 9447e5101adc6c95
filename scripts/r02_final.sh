#!/bin/bash
# final pass of round 2 on one GPU: whole GPU suite, smoke, the driver's bench command, the reference arm
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/f_all_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/f_all_tests.log; tail -n 4 gpurun_out/f_all_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/f_smoke.log
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err ) 2> gpurun_out/f_bench.time
echo "bench rc=$?"; tail -n 3 gpurun_out/f_bench.time
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err ) 2> gpurun_out/f_bench_ref.time
echo "ref rc=$?"; tail -n 3 gpurun_out/f_bench_ref.time
