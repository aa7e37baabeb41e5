"""A/B timing of the persistent flow kernels under debug bits (per-chain durations from CTA 0's layer stamps).
usage: flow_ab.py [batch] bits [bits ...]   (bit 64 is added automatically)"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import maru_b200
from bench import model_files
from maru_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
variants = [int(a) for a in sys.argv[2:]] or [0]
model, _ = model_files('b4c128')
ev = maru_b200.Evaluator(model, gpu=0, max_batch=B, flags=maru_b200.capi.FLAG_NO_GRAPH)
st = synth.random_states(B, 19, seed=7)
dev = torch.device('cuda', 0)
d_states = torch.from_numpy(st).to(dev)
d_out = torch.empty((B, maru_b200.OUTPUT_SIZE), dtype=torch.float32, device=dev)
stream = torch.cuda.ExternalStream(ev.stream(), device=dev)
for bits in variants:
    ev.debug_set(1, bits | 64 | 128)
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
    ev.debug_set(1, bits | (0 if bits & 32 else 128))
    for _ in range(3): ev.forward_states_device(d_states.data_ptr(), d_out.data_ptr(), B)
    ev.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream): e0.record(stream)
    for _ in range(20): ev.forward_states_device(d_states.data_ptr(), d_out.data_ptr(), B)
    with torch.cuda.stream(stream): e1.record(stream)
    ev.sync()
    fwd = e0.elapsed_time(e1) / 20
    print(f'bits {bits:5d}: forward {fwd*1e3:7.1f} us   chains (us, best of 6) ' + ' '.join(f'{c:7.1f}' for c in best) + f'   sum {sum(best):7.1f}', flush=True)
ev.close()
