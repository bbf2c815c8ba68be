#!/bin/bash
# Profiling recipe (run under gpurun on ONE B200; see /opt/skills/guides/B200_PROFILING.md).
# Each ncu command runs only after the same command line exited 0 without ncu (plain run && ncu run).
# Usage: profiles/run_profiles.sh <tag> [fa_stress|seg_counts|fa_train|all]      -> gpurun_out/<tag>_*.{log,csv,ncu-rep}
set -u
TAG=${1:-r01}
ONLY=${2:-all}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"

# ---- fa_stress (BASELINE configs[3], the default bench line): launch list + full capture of the tile engine
if [ $ONLY = all ] || [ $ONLY = fa_stress ]; then
CMD="python bench.py --workload fa_stress --steps 2 --warmup 3 --no-extra"
$CMD > $OUT/${TAG}_fa_stress_plain.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum -k regex:"fa_pos|fa_ref|seg_counts" -c 60 --csv --log-file $OUT/${TAG}_fa_stress_launches.csv $CMD > $OUT/${TAG}_fa_stress_ncu.log 2>&1
$CMD > /dev/null 2>&1 &&
# the dominant kernel: the gradient pass of the two-pass form (fa_pos_grad; fa_pos_tiles* when a fused kernel runs instead)
$NCU --set full --import-source on -k regex:"fa_pos_grad|fa_pos_tiles" -s 3 -c 1 -f -o $OUT/${TAG}_fa_stress_full $CMD > $OUT/${TAG}_fa_stress_full.log 2>&1
# the other kernels of the step: pass A (upper-triangle D tiles -> sign planes) and the near-tie resolve (L2 gather bound)
$NCU --set full --import-source on -k regex:fa_pos_dsign -s 3 -c 1 -f -o $OUT/${TAG}_fa_dsign_full $CMD > $OUT/${TAG}_fa_dsign_full.log 2>&1
$NCU --set full --import-source on -k regex:fa_pos_resolve -s 3 -c 1 -f -o $OUT/${TAG}_fa_resolve_full $CMD > $OUT/${TAG}_fa_resolve_full.log 2>&1
$NCU --set full --import-source on -k regex:fa_pos_pack -s 3 -c 1 -f -o $OUT/${TAG}_fa_pack_full $CMD > $OUT/${TAG}_fa_pack_full.log 2>&1
fi

# ---- seg_counts (BASELINE configs[2])
if [ $ONLY = all ] || [ $ONLY = seg_counts ]; then
CMD="python bench.py --workload seg_counts --steps 2 --warmup 3 --no-extra"
$CMD > $OUT/${TAG}_seg_counts_plain.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum -k regex:"seg_counts" -c 20 --csv --log-file $OUT/${TAG}_seg_counts_launches.csv $CMD > $OUT/${TAG}_seg_counts_ncu.log 2>&1
$CMD > /dev/null 2>&1 &&
$NCU --set full --import-source on -k regex:seg_counts_kernel -s 2 -c 1 -f -o $OUT/${TAG}_seg_counts_full $CMD > $OUT/${TAG}_seg_counts_full.log 2>&1
fi

# ---- fa_train (BASELINE configs[1], reference semantics)
if [ $ONLY = all ] || [ $ONLY = fa_train ]; then
CMD="python bench.py --workload fa_train --steps 3 --warmup 3 --no-extra"
$CMD > $OUT/${TAG}_fa_train_plain.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum -k regex:"fa_ref" -c 40 --csv --log-file $OUT/${TAG}_fa_train_launches.csv $CMD > $OUT/${TAG}_fa_train_ncu.log 2>&1
$CMD > /dev/null 2>&1 &&
$NCU --set full --import-source on -k regex:fa_ref_fused_small -s 4 -c 1 -f -o $OUT/${TAG}_fa_train_full $CMD > $OUT/${TAG}_fa_train_full.log 2>&1
fi
ls -la $OUT | grep $TAG
