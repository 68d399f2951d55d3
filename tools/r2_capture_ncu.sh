#!/bin/bash
# Round-2 evidence, part 1 (run first; tools/r2_summarise.py then refreshes profiles/r2_traffic.json, which bench.py reads
# for roofline.traffic): ncu --set full captures of the step's kernels (bench-sized inputs, 12 of 34 CTU rows) and of the
# frame pass (whole picture).
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'rough_tma|fused_imma|fused_small' -c 12 \
    -f -o gpurun_out/r2_prof python tools/prof_kernels.py 12 > gpurun_out/r2_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'intra_rough|frame_tu_predict|coeff_compact|frame_topk' -s 12 -c 12 \
    -f -o gpurun_out/r2_prof_frame python tools/frame_prof.py --iters 2 > gpurun_out/r2_ncu_frame.log 2>&1
echo captured-ncu
