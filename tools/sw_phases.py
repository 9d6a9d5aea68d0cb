#!/usr/bin/env python
"""Phase timeline (SM cycles) of CTA 0 of the last attn_post_sw_kernel launch (with next-layer fusion) of a forward,
and the average launch time of the node kernel for both tcgen05 variants."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prosim_b200 import lib, synthetic, weights
from prosim_b200.model import ProSimB200
dev = torch.device('cuda', 0)
scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 32
model = ProSimB200(state_dict=weights.random_state_dict(0), device=dev)
pristine = synthetic.clone_batch(synthetic.make_batch(n_scenes=scenes, n_agents=128, n_map=512, steps=20), dev)[0]
ref = None
for mask in (15,):
    lib.set_tensor_core(mask)
    with torch.no_grad():
        for _ in range(2):
            model.forward(synthetic.clone_batch(pristine)[0], 'val')
        lib.profile_enable('attn_post')
        model.forward(synthetic.clone_batch(pristine)[0], 'val')
        ms, n = lib.profile_read()
        lib.profile_enable(None)
        traj = model.forward(synthetic.clone_batch(pristine)[0], 'val')['motion_pred']['_state']['traj'].clone()
    if ref is None:
        ref = traj
    print(f'mask {mask}: attn_post {n} launches, avg {1e3 * ms / n:.1f} us; max |traj - first| {float((traj - ref).abs().max()):.2e}')
lib.set_tensor_core(15)
with torch.no_grad():
    model.forward(synthetic.clone_batch(pristine)[0], 'val')
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 32)()
lib.call('prosim_tc_debug_read', ctypes.cast(buf, ctypes.c_void_p))
t = list(buf)
t0 = t[0]
ne = ['start', 'agg (ffma) done', '-', 'agg->A', 'gate acc', 'u->A', 'out acc', 'xn->A', 'hidden done', 'down acc', 'out,xd->A',
      's acc', 'q->A', 'qhat done']
print('epilogue:', ' | '.join(f'{n} {t[i] - t0}' for i, n in enumerate(ne)))
nm = ['start', 'agg issued', 'a(agg)', 'gate issued', 'a(u)', 'out issued', 'a(xn)', 'ffn issued', 'a(xd)', 'sgq issued', 'a(q)',
      'qhat issued']
print('mma     :', ' | '.join(f'{n} {t[16 + i] - t0}' for i, n in enumerate(nm)))
print('agg detail: kernel entry', t[30] - t0, '| weights half 0 landed', t[2] - t0, '| half 1 waited', t[15] - t0, '| ffma done', t[14] - t0)
print('mma thread waited (cycles): weights', t[28], 'epilogue', t[29], 'timeout id', t[31])
