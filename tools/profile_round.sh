#!/bin/bash
# Round-end measurement pass on the GPU box (run through gpurun): parity tests, the bench line (both arms), the ncu launch list
# of one forward and one `--set full` capture of the four kernels of the first policy a2p / m2p layers.  usage: profile_round.sh r2
set -u
V=${1:-r2}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${V}_gpu_tests.log; cat gpurun_out/${V}_gpu_tests.log
python bench.py > gpurun_out/${V}_bench.json 2> gpurun_out/${V}_bench.err; tail -c 300 gpurun_out/${V}_bench.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${V}_bench_reference.json 2> gpurun_out/${V}_bench_reference.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${V}_launches.csv python tools/profile_forward.py > gpurun_out/${V}_prof.log 2>&1
python tools/launch_summary.py gpurun_out/${V}_launches.csv > gpurun_out/${V}_launches_summary.txt; head -12 gpurun_out/${V}_launches_summary.txt
# encoder (12 layers) + generator (12 layers) = 96 launches of these four kernels come first; then the first tick's a2p / m2p layer
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'attn_edge4|attn_post_sw|edge_qk_kernel|edge_av_kernel' -s 96 -c 8 -f -o gpurun_out/${V}_layer \
    python tools/profile_forward.py --ticks 1 >> gpurun_out/${V}_prof.log 2>&1
ncu -i gpurun_out/${V}_layer.ncu-rep --page raw --csv > gpurun_out/${V}_layer_raw.csv 2>> gpurun_out/${V}_prof.log
python tools/ncu_summary.py gpurun_out/${V}_layer_raw.csv > gpurun_out/${V}_layer_ncu_summary.txt
python tools/kernel_traffic.py gpurun_out/${V}_layer_raw.csv profiles/${V}_layer_ncu_summary.txt > gpurun_out/${V}_kernel_traffic.json
ncu -i gpurun_out/${V}_layer.ncu-rep --page source --csv -k regex:attn_post_sw -c 1 > gpurun_out/${V}_post_sw_src.csv 2>> gpurun_out/${V}_prof.log
# single scene (the literal BASELINE configs[2]): launch list
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${V}_launches_1scene.csv python tools/profile_forward.py --scenes 1 >> gpurun_out/${V}_prof.log 2>&1
python tools/launch_summary.py gpurun_out/${V}_launches_1scene.csv > gpurun_out/${V}_launches_1scene_summary.txt
python tools/sw_phases.py > gpurun_out/${V}_sw_phases.txt 2>&1
python tools/parity_sweep.py 0 9 15 31 > gpurun_out/${V}_parity_sweep.txt 2>&1
ls -la gpurun_out/${V}_*
