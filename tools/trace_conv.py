"""Pipeline trace of one conv launch (clock64 stamps per tile and warp role)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import maru_b200
from bench import model_files
from maru_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
which = int(sys.argv[2]) if len(sys.argv) > 2 else 2      # trunk launch index: 0 stem, 1 reduce, 2 c1, 3 c2 ...
model, _ = model_files('b4c128')
ev = maru_b200.Evaluator(model, gpu=0, max_batch=B, flags=1)
st = synth.random_states(B, 19, seed=7)
for _ in range(3): ev.forward_states(st)
ev.debug_set(2, which)
ev.forward_states(st)
tr = ev.debug_trace(148)
names = ['prod_wait', 'prod_issue', 'mma_wait', 'mma_ready', 'mma_issued', 'epi_wait', 'epi_ready', 'epi_done']
for cta in (0, 73, 147):
    t = tr[cta]
    t0 = t[0, 0] if t[0, 0] else t[0, 2]
    print(f'--- CTA {cta} (cycles relative to first producer wait)')
    for i in range(8):
        if t[i, 2] == 0: break
        print(f'tile {i}: ' + '  '.join(f'{n}={int(t[i, k] - t0):6d}' for k, n in enumerate(names)))
# summary over CTAs: per-tile period and wait components
per, w_full, w_tfull, epi = [], [], [], []
for cta in range(148):
    t = tr[cta]
    n = int((t[:, 4] != 0).sum())
    if n >= 4:
        per.append((t[n - 1, 4] - t[1, 4]) / (n - 2))
        w_full.append(np.mean(t[1:n, 3] - t[1:n, 2]))
        w_tfull.append(np.mean(t[1:n, 6] - t[1:n, 5]))
        epi.append(np.mean(t[1:n, 7] - t[1:n, 6]))
print('mean cycles: tile period %.0f, MMA-warp wait %.0f, epilogue wait %.0f, epilogue work %.0f' %
      (np.mean(per), np.mean(w_full), np.mean(w_tfull), np.mean(epi)))
ev.close()
