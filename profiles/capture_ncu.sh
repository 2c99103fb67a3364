#!/bin/sh
# ncu evidence of the float32-faithful tensor-core mode (run on the GPU box through gpurun):
#   sh profiles/capture_ncu.sh <tag>      -> gpurun_out/<tag>_*.csv (+ one .ncu-rep of the 7^3 stem for source-level reading)
# The big .ncu-rep files are summarised on the box (profiles/summarize_ncu.py) and deleted: gpurun_out/ is capped at 64 MiB.
TAG="${1:-r02}"
OUT=gpurun_out
mkdir -p $OUT
# 1. every launch of one bench step with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${TAG}_ncu_launches_bench.csv python profiles/run_step.py > /dev/null 2>&1
# 2. ncu --set full of the pose net's convolutions (16 cubes per launch keeps the replay memory small), of the 7^3 stem at
#    the bench's 80 cubes per launch, and of the un-projection / max-pool / NMS / soft-argmax kernels of a bench step
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc -c 40 \
    -o $OUT/${TAG}_ncu_pose_v2v_16 python profiles/pose_v2v_only.py --cubes 16 > /dev/null 2>&1
python profiles/summarize_ncu.py $OUT/${TAG}_ncu_pose_v2v_16.ncu-rep > $OUT/${TAG}_ncu_pose_v2v_16cubes_summary.csv
rm -f $OUT/${TAG}_ncu_pose_v2v_16.ncu-rep
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc -c 1 \
    -o $OUT/${TAG}_ncu_stem_80 python profiles/pose_v2v_only.py --layers stem > /dev/null 2>&1
python profiles/summarize_ncu.py $OUT/${TAG}_ncu_stem_80.ncu-rep > $OUT/${TAG}_ncu_stem_80cubes_summary.csv
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:unproject|maxpool|softargmax|nms" -c 12 \
    -o $OUT/${TAG}_ncu_k1_k3_k4 python profiles/run_step.py > /dev/null 2>&1
python profiles/summarize_ncu.py $OUT/${TAG}_ncu_k1_k3_k4.ncu-rep > $OUT/${TAG}_ncu_k1_k3_k4_summary.csv
rm -f $OUT/${TAG}_ncu_k1_k3_k4.ncu-rep
ls -la $OUT
