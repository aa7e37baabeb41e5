"""Layer-by-layer timeline of one graph-replayed forward (globaltimer stamps of the conv launches)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import maru_b200
from bench import model_files
from maru_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model, _ = model_files('b4c128')
ev = maru_b200.Evaluator(model, gpu=0, max_batch=B, flags=int(sys.argv[2]) if len(sys.argv) > 2 else 0)
st = synth.random_states(B, 19, seed=7)
ev.debug_set(3, 1)
if len(sys.argv) > 3: ev.debug_set(1, int(sys.argv[3]))
for _ in range(4): ev.forward_states(st)
tl = ev.debug_timeline()
t0 = tl[0, 0, 0]
print('seq  entry  prologue  dep_ok  operands  first_acc  exit | lastCTA: entry ... exit     (us since first conv entry)')
prev_exit = t0
for i in range(64):
    if tl[i, 0, 0] == 0: continue
    a = (tl[i, 0, :6] - t0) / 1e3
    b = (tl[i, 1, :6] - t0) / 1e3
    dealloc = (tl[i, 0, 6] - tl[i, 0, 5]) / 1e3 if tl[i, 0, 6] else float('nan')
    print(f'{i:3d} {a[0]:7.1f} {a[1]:7.1f} {a[2]:7.1f} {a[3]:7.1f} {a[4]:7.1f} {a[5]:7.1f} | {b[0]:7.1f} {b[2]:7.1f} {b[5]:7.1f}   dur {a[5]-a[0]:5.1f}  gap-from-prev-exit {a[0]-prev_exit:5.1f}  tmem-dealloc {dealloc:5.2f} (fence {(tl[i,0,7]-tl[i,0,5])/1e3:5.2f})')
    prev_exit = max(a[5], b[5])
ev.close()
