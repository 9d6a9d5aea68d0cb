#!/bin/bash
# Round-end measurement pass on the GPU box (run through gpurun): parity tests, the bench line, the ncu launch list of one
# forward and one `--set full` capture of the four kernels of the first policy a2p / m2p layers.  usage: profile_round.sh v8
set -u
V=${1:-v8}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/gpu_tests_$V.log; cat gpurun_out/gpu_tests_$V.log
python bench.py > gpurun_out/bench_$V.json 2> gpurun_out/bench_$V.err; tail -c 600 gpurun_out/bench_$V.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r1_launches_$V.csv python tools/profile_forward.py > gpurun_out/prof_$V.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'attn_edge4|attn_post_tc|edge_qk|edge_av' -s 96 -c 8 -f -o gpurun_out/r1_layer_$V \
    python tools/profile_forward.py --ticks 1 >> gpurun_out/prof_$V.log 2>&1
ncu -i gpurun_out/r1_layer_$V.ncu-rep --page raw --csv > gpurun_out/r1_layer_${V}_raw.csv 2>> gpurun_out/prof_$V.log
ncu -i gpurun_out/r1_layer_$V.ncu-rep --page source --csv -k regex:attn_edge4 -c 1 > gpurun_out/edge4_src_$V.csv 2>> gpurun_out/prof_$V.log
ls -la gpurun_out/*$V*
# the tensor-core PointNet / K'V' kernels (first launches of the forward: map polylines, agent histories, one K'V' launch)
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'pointnet_tc|attn_kv_tc' -c 4 -f -o gpurun_out/r1_pntc_$V \
    python tools/profile_forward.py --ticks 1 >> gpurun_out/prof_$V.log 2>&1
ncu -i gpurun_out/r1_pntc_$V.ncu-rep --page raw --csv > gpurun_out/r1_pntc_${V}_raw.csv 2>> gpurun_out/prof_$V.log
ls -la gpurun_out/*pntc*$V*
