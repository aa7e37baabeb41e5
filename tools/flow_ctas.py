"""Entry / start / finish / exit of every CTA of one persistent flow kernel launch (conv_flow.cu)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import maru_b200
from bench import model_files
from maru_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
chain = int(sys.argv[2]) if len(sys.argv) > 2 else 0
model, _ = model_files('b4c128')
ev = maru_b200.Evaluator(model, gpu=0, max_batch=B)
st = synth.random_states(B, 19, seed=7)
ev.debug_set(1, 128)
ev.debug_set(4, 1000)      # trace buffer on, no per-item CTA
ev.debug_set(5, chain)
for _ in range(3): ev.forward_states(st)
t = ev.debug_trace(72).reshape(-1, 8)[1024:1024 + 148].astype(np.float64)
t0 = t[:, 0].min()
e, s, f, x, k = [(t[:, i] - t0) / 1e3 for i in range(4)] + [t[:, 4]]
print(f'148 CTAs: entry {e.min():.1f}..{e.max():.1f}  start {s.min():.1f}..{s.max():.1f}  finish {f.min():.1f}..{f.max():.1f}  exit {x.min():.1f}..{x.max():.1f} us')
for kk in sorted(set(k.astype(int))):
    m = k == kk
    print(f'  k={kk}: {m.sum():3d} CTAs, finish {f[m].min():.1f}..{f[m].max():.1f} (mean {f[m].mean():.1f}), busy {(f[m]-s[m]).mean():.1f} us')
print('cta: ' + ' '.join(f'{i}:{f[i]:.0f}' for i in range(0, 148, 6)))
ev.close()
