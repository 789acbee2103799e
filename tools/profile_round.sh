#!/bin/bash
# One GPU: the ncu evidence of a round (run under gpurun): launch list of the bench command, and one
# `--set full` capture of each hot kernel of a 16384^2 step. Outputs under gpurun_out/<tag>_*.
tag=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches_bench_16384.csv \
    python bench.py --steps 2 --warmup 3 --cpu-steps 0 --eager-steps 0 --secondary-n 0 > gpurun_out/${tag}_launches_bench.json 2> gpurun_out/${tag}_launches_bench.err
NSB_CHAIN_PIECES=1 ncu --set full --clock-control none --import-source on -k regex:"k_alpha1024|k_chain_B_tma|k_chain_G2|k_chain_F_tma" -s 5 -c 4 \
    -o gpurun_out/${tag}_full_16384 python tools/one_step.py 16384 1 > gpurun_out/${tag}_full.log 2>&1
tail -2 gpurun_out/${tag}_full.log
