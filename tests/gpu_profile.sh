#!/bin/bash
# gpu_profile.sh <tag> -- the ncu evidence of a build (run under gpurun, one GPU): launch list with light metrics for every
# kernel of one V1 sign + verify step of 2^19 items, and a --set full capture (with SASS-level source counters) of the five
# big kernels at 2^18 items, exported to csv on the box (the .ncu-rep itself is too large to bring back every time).
tag=${1:-r02}
mkdir -p gpurun_out
export PLUME_DEVICE_SPLIT=0   # one kernel sequence per batch: every stage kernel is profiled whole
M=gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sass__inst_executed_local_loads,sass__inst_executed_local_stores
ncu --metrics $M --clock-control none -k regex:'^k_' -s 45 -c 15 --csv --log-file gpurun_out/${tag}_all.csv \
    python bench.py --steps 1 --warmup 3 --log2-batch 19 --no-cpu-baseline > gpurun_out/${tag}_all.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_sign_comb_tab|k_sign_comb_lad|k_verify_mul_a|k_verify_lad_b|k_verify_tab_b' -s 15 -c 5 \
    -f -o gpurun_out/${tag}_full python bench.py --steps 1 --warmup 3 --log2-batch 18 --no-cpu-baseline > gpurun_out/${tag}_full.log 2>&1
ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_full.ncu-rep --page source --csv > gpurun_out/${tag}_full_source.csv 2>/dev/null
ls -la gpurun_out/${tag}_full*
# keep the report only when it fits the 64 MiB that come back
if [ $(stat -c %s gpurun_out/${tag}_full.ncu-rep) -gt 40000000 ]; then rm -f gpurun_out/${tag}_full.ncu-rep; fi
