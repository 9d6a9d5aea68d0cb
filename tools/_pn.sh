ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:pointnet -c 3 -f -o gpurun_out/pn_v2 python tools/profile_forward.py --ticks 2 > gpurun_out/pn_prof.log 2>&1
ncu -i gpurun_out/pn_v2.ncu-rep --page raw --csv > gpurun_out/pn_v2_raw.csv
PROSIM_POINTNET_V1=1 ncu --profile-from-start off --set full --clock-control none -k regex:pointnet -c 3 -f -o gpurun_out/pn_v1 python tools/profile_forward.py --ticks 2 >> gpurun_out/pn_prof.log 2>&1
ncu -i gpurun_out/pn_v1.ncu-rep --page raw --csv > gpurun_out/pn_v1_raw.csv
python tools/ncu_summary.py gpurun_out/pn_v2_raw.csv > gpurun_out/pn_v2_summary.txt
python tools/ncu_summary.py gpurun_out/pn_v1_raw.csv > gpurun_out/pn_v1_summary.txt
grep -h "Kernel Name\|time_duration" gpurun_out/pn_v2_summary.txt gpurun_out/pn_v1_summary.txt
