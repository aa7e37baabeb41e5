"""Per-layer timeline of the persistent flow kernels (CTA 0 %globaltimer stamps, debug bit 6)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import maru_b200
from bench import model_files
from maru_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
name = sys.argv[2] if len(sys.argv) > 2 else 'b4c128'
model, _ = model_files(name)
ev = maru_b200.Evaluator(model, gpu=0, max_batch=B, flags=maru_b200.capi.FLAG_NO_GRAPH)
st = synth.random_states(B, 19, seed=7)
ev.debug_set(1, 64 | 128 | (int(sys.argv[3]) if len(sys.argv) > 3 else 0))
for _ in range(3): ev.forward_states(st)
s = ev.debug_stamps().astype(np.float64).reshape(-1, 4)
t0 = s[0, 0]
print('row  layer_start  weights_landed  cta0_done   (us since the first layer of the first chain; dt = to next row)')
for i in range(len(s) - 1):
    if s[i, 0] == 0: continue
    a = (s[i] - t0) / 1e3
    nxt = (s[i + 1, 0] - t0) / 1e3 if s[i + 1, 0] else float('nan')
    if s[i, 2] == 0:
        print(f'{i:3d} {a[0]:9.1f}   -- end of chain --')
    else:
        print(f'{i:3d} {a[0]:9.1f} {a[1]:9.1f} {a[2]:9.1f}   dt {nxt - a[0]:6.1f}')
ev.close()
