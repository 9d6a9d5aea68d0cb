python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/gpu_tests_sm.log; cat gpurun_out/gpu_tests_sm.log
python tools/kernel_breakdown.py 2>&1 | tail -3 > gpurun_out/breakdown_sm.txt
python tools/split_timeline.py 1 400 30 > gpurun_out/split_timeline1.txt 2>&1
python tools/split_sweep.py > gpurun_out/split_sweep3.txt 2>&1
python - <<PY
import json
for l in open("gpurun_out/breakdown_sm.txt"):
    if l.startswith("{\"forward_ms\""):
        d=json.loads(l); print(d["forward_ms"], d["attn_edge"], d["attn_post"], d["edge_qk"], d["edge_av"])
PY
grep parts gpurun_out/split_timeline1.txt; cat gpurun_out/split_sweep3.txt | tail -5
