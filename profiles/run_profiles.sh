#!/bin/bash
# Profiling recipe (run under gpurun on ONE B200; see /opt/skills/guides/B200_PROFILING.md).
# Each ncu command runs only after the same command line exited 0 without ncu.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"

# 1. launch lists (cold-cache, serialised: compare SHARES)
CMD_FA="python bench.py --workload fa_train --steps 3 --warmup 3 --no-extra"
$CMD_FA > $OUT/${TAG}_fa_train_plain.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/${TAG}_fa_train_launches.csv $CMD_FA > $OUT/${TAG}_fa_train_ncu.log 2>&1

CMD_SEG="python bench.py --workload seg_counts --steps 2 --warmup 3 --no-extra"
$CMD_SEG > $OUT/${TAG}_seg_counts_plain.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum -k regex:"seg_counts|memset|Memset" -c 40 --csv --log-file $OUT/${TAG}_seg_counts_launches.csv $CMD_SEG > $OUT/${TAG}_seg_counts_ncu.log 2>&1

# 2. full captures of the top kernels
$NCU --set full --import-source on -k regex:seg_counts_kernel -s 2 -c 2 -f -o $OUT/${TAG}_seg_counts_full $CMD_SEG > $OUT/${TAG}_seg_counts_full.log 2>&1
$NCU --set full --import-source on -k regex:fa_ref -s 8 -c 8 -f -o $OUT/${TAG}_fa_train_full $CMD_FA > $OUT/${TAG}_fa_train_full.log 2>&1
ls -la $OUT
