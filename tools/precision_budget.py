"""CPU emulation of 16-bit operand / storage choices for EDSR x4 (16 blocks, he_normal random init), against the
float64 oracle.  Shows which roundings decide the SR-output max-abs error; DESIGN.md "Precision" quotes this table.

    python tools/precision_budget.py
"""
import sys, numpy as np, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'super-resolution-images-for-3d-printing-defect-detection_b200'))
from oracle import convnets as oc
from srb200 import weights, synth
import torch.nn.functional as F
torch.set_num_threads(8)
def rnd(t, kind):
    if kind=='bf16': return t.bfloat16().float()
    if kind=='fp16': return t.half().float()
    if kind=='fp16+e5m2':      # hi = fp16(t), lo = e5m2(t - hi): the two-tensor trunk of precision="fp16", trunk="pair8"
        hi = t.half().float()
        return hi + (t - hi).to(torch.float8_e5m2).float()
    return t
def conv(x, w, name, wk):
    k = torch.from_numpy(w[name+'/kernel']).permute(3,2,0,1).contiguous()
    return F.conv2d(x, rnd(k,wk), torch.from_numpy(w[name+'/bias']), padding=1)
def edsr(w, x, nb, wk_body, ak_body, trunk, wk_late, ak_late):
    x = torch.from_numpy(x).permute(0,3,1,2).contiguous()
    h = conv(x, w, 'head', None)          # head conv fp32 weights
    head = rnd(h, trunk)
    h = head
    for i in range(nb):
        t = rnd(F.relu(conv(rnd(h,ak_body), w, f'rb{i}_c1', wk_body)), ak_body)   # (rnd(h, fp16) of an fp16+e5m2 trunk = its hi half)
        h = rnd(h + 0.1*conv(t, w, f'rb{i}_c2', wk_body), trunk)
    if trunk=='fp16+e5m2': h = h.half().float()      # the next conv's operand is the hi half alone
    h = rnd(conv(rnd(h,ak_late), w, 'body', wk_late) + head, ak_late)
    h = rnd(oc.depth_to_space(conv(h, w, 'up0', wk_late),2), ak_late)
    h = rnd(oc.depth_to_space(conv(h, w, 'up1', wk_late),2), ak_late)
    h = conv(h, w, 'tail', wk_late)
    return h.permute(0,2,3,1).numpy()
w = weights.edsr_weights(4)
lr = synth.area_downsample(synth.hr_batch(2,192,192),4)
ref = oc.edsr_forward(w, lr, 4, 16, dtype=torch.float64)
def report(tag, **kw):
    out = edsr(w, lr, 16, **kw)
    pre = out
    e = np.abs(np.clip(out,0,1)-ref)
    print(f"{tag:50s} max {e.max():.4f}  mean {e.mean():.5f}  pre-clip std {pre.std():.2f}")
report('fp32 emul', wk_body=None, ak_body=None, trunk=None, wk_late=None, ak_late=None)
report('all bf16 (trunk bf16)', wk_body='bf16', ak_body='bf16', trunk='bf16', wk_late='bf16', ak_late='bf16')
report('bf16, fp32 trunk', wk_body='bf16', ak_body='bf16', trunk=None, wk_late='bf16', ak_late='bf16')
report('bf16 body + fp32 trunk + fp16 late', wk_body='bf16', ak_body='bf16', trunk=None, wk_late='fp16', ak_late='fp16')
report('bf16 body + bf16 trunk + fp16 late', wk_body='bf16', ak_body='bf16', trunk='bf16', wk_late='fp16', ak_late='fp16')
report('all fp16 (trunk fp16)', wk_body='fp16', ak_body='fp16', trunk='fp16', wk_late='fp16', ak_late='fp16')
report('fp16, fp32 trunk', wk_body='fp16', ak_body='fp16', trunk=None, wk_late='fp16', ak_late='fp16')
report('weights bf16 only', wk_body='bf16', ak_body=None, trunk=None, wk_late='bf16', ak_late=None)
report('acts bf16 only', wk_body=None, ak_body='bf16', trunk='bf16', wk_late=None, ak_late='bf16')
report('fp16, fp16+e5m2 trunk (pair8)', wk_body='fp16', ak_body='fp16', trunk='fp16+e5m2', wk_late='fp16', ak_late='fp16')
