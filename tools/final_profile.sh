set -u
V=r2f
timeout 100 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${V}_launches.csv python tools/profile_forward.py > gpurun_out/${V}_prof.log 2>&1
python tools/launch_summary.py gpurun_out/${V}_launches.csv > gpurun_out/${V}_launches_summary.txt; head -8 gpurun_out/${V}_launches_summary.txt
timeout 70 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${V}_launches_1scene.csv python tools/profile_forward.py --scenes 1 >> gpurun_out/${V}_prof.log 2>&1
python tools/launch_summary.py gpurun_out/${V}_launches_1scene.csv > gpurun_out/${V}_launches_1scene_summary.txt; head -6 gpurun_out/${V}_launches_1scene_summary.txt
timeout 150 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'attn_edge4|attn_post_sw|edge_qk_kernel|edge_av_kernel' -s 110 -c 8 -f -o gpurun_out/${V}_layer python tools/profile_forward.py --ticks 1 >> gpurun_out/${V}_prof.log 2>&1
ncu -i gpurun_out/${V}_layer.ncu-rep --page raw --csv > gpurun_out/${V}_layer_raw.csv 2>> gpurun_out/${V}_prof.log
python tools/ncu_summary.py gpurun_out/${V}_layer_raw.csv > gpurun_out/${V}_layer_ncu_summary.txt; head -5 gpurun_out/${V}_layer_ncu_summary.txt
python tools/sw_phases.py > gpurun_out/${V}_sw_phases.txt 2>&1; tail -4 gpurun_out/${V}_sw_phases.txt | cut -c1-200
rm -f gpurun_out/${V}_layer.ncu-rep
