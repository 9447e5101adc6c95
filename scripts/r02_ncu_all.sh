#!/bin/bash
# ncu evidence of round 2 (one call): C5 k_spw with / without the L2 persisting window, C3 .dc generated kernel, C2 generated
# kernel with instance-major / parameter-major parameters (L1 sectors per request), launch list of the default bench line
mkdir -p gpurun_out
C5="python bench.py --workload c5 --scale 0.09 --tran-stop 3e-9 --steps 1 --warmup 3 --no-cpu"
$C5 > gpurun_out/u_c5_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_spw -s 6 -c 1 -o gpurun_out/u_c5_persist $C5 > gpurun_out/u_c5_f1.log 2>&1
ncu --set full --clock-control none -k regex:k_spw -s 6 -c 1 -o gpurun_out/u_c5_nopersist $C5 --opt l2_persist=0 > gpurun_out/u_c5_f0.log 2>&1
C3="python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu"
$C3 > gpurun_out/u_c3_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:ftsb_jit -s 4 -c 1 -o gpurun_out/u_c3 $C3 > gpurun_out/u_c3_f.log 2>&1
python scripts/soa_probe.py 0 > gpurun_out/u_soa_plain.log 2>&1 &&
ncu --metrics l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:ftsb_jit --csv --log-file gpurun_out/u_soa_inst.csv python scripts/soa_probe.py 0 > gpurun_out/u_soa0.log 2>&1
ncu --metrics l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:ftsb_jit --csv --log-file gpurun_out/u_soa_param.csv python scripts/soa_probe.py 1 > gpurun_out/u_soa1.log 2>&1
ALL="python bench.py --steps 2 --warmup 3 --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/u_launches_all.csv $ALL > gpurun_out/u_all_l.log 2>&1
ls -la gpurun_out/u_*
