"""Per-layer timeline of the persistent chain kernels (CTA 0 %globaltimer stamps, debug bit 6)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import maru_b200
from bench import model_files
from maru_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model, _ = model_files(sys.argv[2] if len(sys.argv) > 2 else 'b4c128')
ev = maru_b200.Evaluator(model, gpu=0, max_batch=B, flags=int(sys.argv[3]) if len(sys.argv) > 3 else 0)
st = synth.random_states(B, 19, seed=7)
ev.debug_set(1, 32 | 64)
for _ in range(4):
    ev.forward_states(st)
s = ev.debug_stamps().astype(np.float64)
t0 = s[0, 0]
print('row  start  weights  operands  done   | layer us  start->operands  operands->done  done->next start')
for i in range(127):
    if s[i, 0] == 0:
        break
    a = (s[i] - t0) / 1e3
    nxt = (s[i + 1, 0] - t0) / 1e3 if s[i + 1, 0] else float('nan')
    if s[i, 1] == 0 and s[i, 2] == 0:
        print(f'{i:3d} {a[0]:8.1f}   -- end of chain --')
        continue
    print(f'{i:3d} {a[0]:8.1f} {a[3]:8.1f} {a[1]:8.1f} {a[2]:8.1f} | {nxt - a[0]:6.1f} {a[1] - a[0]:8.1f} {a[2] - a[1]:8.1f} {nxt - a[2]:8.1f}')
ev.close()
# per-tile trace of CTA 0 in one layer of the first chain (argv[4] = layer index)
if len(sys.argv) > 4:
    ev = maru_b200.Evaluator(model, gpu=0, max_batch=B, flags=0)
    ev.debug_set(1, 32 | 64)
    ev.debug_set(2, int(sys.argv[4]))
    for _ in range(3):
        ev.forward_states(st)
    raw = ev.debug_stamps().reshape(-1)
    tr = raw[256:256 + 32 * 8].reshape(32, 8).astype(np.float64)
    base = raw[4 * int(sys.argv[4])]
    print('tile  prod_issue  operands  mma_issued  acc_ready  epi_done   (us since layer start)')
    for i in range(32):
        if tr[i, 0] == 0 and tr[i, 1] == 0:
            break
        print(f'{i:3d} ' + ' '.join(f'{(v - base) / 1e3:10.2f}' if v else '        --' for v in tr[i, :5]))
    ev.close()
