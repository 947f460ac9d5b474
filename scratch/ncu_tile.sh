#!/bin/bash
# usage: ncu_tile.sh <tag>   (env selects the variant)
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__warps_active.avg.per_cycle_active,smsp__inst_executed_op_shared_ld.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum
ncu --metrics $M --clock-control none -k regex:leaf_topk_tile -s 8 -c 1 --csv --log-file gpurun_out/tile_$1.csv python scratch/exp1.py > gpurun_out/tile_$1.log 2>&1
