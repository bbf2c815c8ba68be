#!/bin/bash
# ncu capture of the FA(position) tile engine on one sample of BASELINE configs[3] (128x256 positions), C per branch = $1
C=${1:-128}; TAG=${2:-r01f}     # $3 = batch (default 1), $4 = precision (tf32 | f16 | fp32)
CMD="python tools/perf_pos.py ${3:-1},$C,128,256 ${4:-tf32}"
$CMD > gpurun_out/${TAG}_pos_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fa_pos_tiles -s 2 -c 1 -f -o gpurun_out/${TAG}_pos_c$C $CMD > gpurun_out/${TAG}_pos_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_pos_plain.log; tail -2 gpurun_out/${TAG}_pos_ncu.log
