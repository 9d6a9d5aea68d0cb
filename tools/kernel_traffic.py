#!/usr/bin/env python
"""profiles/r<N>_kernel_traffic.json from the raw page of the per-layer `ncu --set full` capture (tools/profile_round.sh):
DRAM bytes, duration and tensor-pipe activity of the four kernels of the first policy a2p layer and of the first m2p layer.
usage: kernel_traffic.py gpurun_out/r1_layer_v8_raw.csv profiles/r1_layer_v8_ncu_summary.txt > profiles/r1_kernel_traffic.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}


def val(d, name):
    v, u = float(d[idx[name]].replace(',', '')), units[idx[name]]
    scale = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1.0, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'msecond': 1e3,
             'usecond': 1.0, 'nsecond': 1e-3}.get(u, 1.0)
    return v * scale


labels = ['policy a2p layer 0 (380k edges, 4096 rows)'] * 4 + ['policy m2p layer 0 (163k edges, 4096 rows)'] * 4
out = {}
for d, lab in zip(data, labels):
    name = d[idx['Kernel Name']].split('(')[0].replace('void ', '').replace('prosim::', '').split('<')[0]
    if name == 'attn_post_sw_kernel':
        name = 'psw::attn_post_sw_kernel'
    out[f'{name} {"a2p" if "a2p" in lab else "m2p"}'] = {
        'launch': lab,
        'dram_bytes_read': val(d, 'dram__bytes_read.sum'),
        'dram_bytes_write': val(d, 'dram__bytes_write.sum'),
        'duration_us_under_ncu': val(d, 'gpu__time_duration.sum'),
        'tensor_pipe_active_pct': float(d[idx['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']]),
        'source': f'{sys.argv[2]} (ncu --set full --clock-control none)',
    }
print(json.dumps(out, indent=1))
