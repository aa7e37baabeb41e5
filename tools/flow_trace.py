"""Per-item timeline of one CTA of a persistent flow kernel (conv_flow.cu): where does a tile's time go?
usage: flow_trace.py [batch] [cta] [flow launch] [debug bits]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import maru_b200
from bench import model_files
from maru_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cta = int(sys.argv[2]) if len(sys.argv) > 2 else 0
chain = int(sys.argv[3]) if len(sys.argv) > 3 else 0
bits = int(sys.argv[4]) if len(sys.argv) > 4 else 0
model, _ = model_files('b4c128')
ev = maru_b200.Evaluator(model, gpu=0, max_batch=B, flags=maru_b200.capi.FLAG_NO_GRAPH)
st = synth.random_states(B, 19, seed=7)
ev.debug_set(1, bits | 128)
ev.debug_set(4, cta)
ev.debug_set(5, chain)
for _ in range(3): ev.forward_states(st)
t = ev.debug_trace(16).reshape(512, 8).astype(np.float64)
t0 = t[0, 0]
print('item  deps_met  requested  landed  mma_issued  acc_ready  stored  published | mma period, ready->published')
prev = None
for g in range(256):
    if t[g, 0] == 0: break
    a = (t[g, :7] - t0) / 1e3
    per = a[3] - prev if prev is not None else float('nan')
    prev = a[3]
    dw = (t[g, 0] - t[g, 7]) / 1e3 if t[g, 7] else 0.0
    print(f'{g:3d} {a[0]:8.2f} {a[1]:8.2f} {a[2]:8.2f} {a[3]:8.2f} {a[4]:8.2f} {a[5]:8.2f} {a[6]:8.2f}   | {per:6.2f} {a[6]-a[4]:6.2f}  dep-wait {dw:5.2f}')
print('publisher per item: raw count seen, after acquire fence, after proxy fence | producer: polls done, (fence), ring waits done, slot written, (copies issued) (us)')
for g in range(40):
    if t[320 + g, 0] == 0: break
    a = (t[320 + g, :6] - t0) / 1e3
    print(f'item {g:2d} {a[0]:8.2f} {a[1]:8.2f} {a[2]:8.2f} | {(t[g,7]-t0)/1e3:8.2f} {a[3]:8.2f} {(t[g,0]-t0)/1e3:8.2f} {a[4]:8.2f} {a[5]:8.2f} {(t[g,1]-t0)/1e3:8.2f}')
print('aux warp per layer: weights requested / first-tile flag / last-tile flag published')
for l in range(64):
    if t[256 + l, 0] == 0: break
    a = (t[256 + l, :3] - t0) / 1e3
    print(f'layer {l:2d} {a[0]:8.2f} {a[1]:8.2f} {a[2]:8.2f}')
ev.close()
