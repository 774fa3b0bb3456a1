#!/bin/bash
# Round-2 capture protocol (run on the GPU box through gpurun): one `ncu --set full` capture of every kernel of the first
# device-timed step (after its L2 flush) and the launch list of the same command.
#   bash profiles/r02_capture.sh <tag>
tag=${1:-r2}
K='regex:step_kernel|dbscan_big_kernel|pose_feature_kernel|conv_slab_kernel|gemm_tc'
# 12 priming + 3 warm-up steps x 7 kernels = 105 launches of the library before the first timed step
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 105 --launch-count 7 \
    -o gpurun_out/${tag}_full -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1
echo capture rc=$?
