"""Bring-up: flow kernel vs per-layer launches on a list of (model, batch) cases; prints max errors."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import maru_b200
from bench import model_files
from maru_b200 import synth
cases = [a.split(':') for a in sys.argv[1:]] or [['b2c64', '1']]
for name, n in cases:
    n = int(n)
    model, _ = model_files(name)
    ev = maru_b200.Evaluator(model, gpu=0, max_batch=max(n, 16), flags=maru_b200.capi.FLAG_NO_GRAPH)
    st = synth.random_states(n, 19, seed=5 + n)
    ev.debug_set(1, 32)
    want = ev.forward_states(st)
    l0 = ev.info().launches_per_forward
    ev.debug_set(1, 128)
    print(name, n, 'per-layer ok, launches', l0, flush=True)
    got = ev.forward_states(st)
    print(name, n, 'flow launches', ev.info().launches_per_forward, 'max err', float(np.abs(got - want).max()), flush=True)
    ev.close()
