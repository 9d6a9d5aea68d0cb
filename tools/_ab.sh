python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/kernel_breakdown.py 2>&1 | tail -3 > gpurun_out/breakdown_kvvec.txt
python - <<PY
import json
for l in open("gpurun_out/breakdown_kvvec.txt"):
    if l.startswith("{\"forward_ms\""):
        d=json.loads(l); print(round(d["forward_ms"],2), "pointnet", d["pointnet"], "kv", d["attn_kv"])
PY
