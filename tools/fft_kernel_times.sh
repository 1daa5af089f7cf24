#!/bin/bash
# per-kernel durations and pipe utilisation of the hand-written solve (ncu metrics pass; not a bench number)
# usage: tools/fft_kernel_times.sh <tag>
tag=$1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum
PM_SOLVE_MODE=fft2_split PM_REPS=3 timeout 300 ncu --metrics $M --clock-control none -k regex:"fft2d_kernel|xsolve2_kernel" -s 5 -c 5 --csv --log-file gpurun_out/${tag}_split.csv python tools/run_solve_once.py > /dev/null 2>&1
PM_SOLVE_MODE=fft2_l2 PM_REPS=3 timeout 300 ncu --metrics $M --clock-control none -k regex:"fft2d_kernel|xsolve2_kernel" -s 3 -c 3 --csv --log-file gpurun_out/${tag}_l2.csv python tools/run_solve_once.py > /dev/null 2>&1
python tools/ncu_table.py gpurun_out/${tag}_split.csv
python tools/ncu_table.py gpurun_out/${tag}_l2.csv
