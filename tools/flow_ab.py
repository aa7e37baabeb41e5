"""A/B timing of the persistent flow kernels under debug bits (per-chain durations from CTA 0's layer stamps).
usage: flow_ab.py [batch] bits [bits ...]   (bit 64 is added automatically)"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import maru_b200
from bench import model_files
from maru_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
variants = [int(a) for a in sys.argv[2:]] or [0]
model, _ = model_files('b4c128')
ev = maru_b200.Evaluator(model, gpu=0, max_batch=B, flags=maru_b200.capi.FLAG_NO_GRAPH)
st = synth.random_states(B, 19, seed=7)
for bits in variants:
    ev.debug_set(1, bits | 64)
    best = None
    for _ in range(6):
        ev.forward_states(st)
        s = ev.debug_stamps().astype(np.float64).reshape(-1, 4)
        chains, start = [], None
        for i in range(len(s)):
            if s[i, 0] == 0: break
            if start is None: start = s[i, 0]
            if s[i, 2] == 0:
                chains.append((s[i, 0] - start) / 1e3); start = None
        best = chains if best is None else [min(a, b) for a, b in zip(best, chains)]
    print(f'bits {bits:5d}: chains (us, best of 6) ' + ' '.join(f'{c:7.1f}' for c in best) + f'   sum {sum(best):7.1f}', flush=True)
ev.close()
