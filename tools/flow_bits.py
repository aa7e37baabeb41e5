"""Forward time of b4c128 under the flow kernel's timing-experiment bits (EXPERIMENTS builds; results are wrong with any bit).
usage: flow_bits.py [batch] bits [bits ...]     4|16 no dependency waits, 8 no slab loads, 32 no MMAs, 128 no epilogue work"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import maru_b200
from bench import model_files
from maru_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
variants = [int(a) for a in sys.argv[2:]] or [0]
model, _ = model_files('b4c128')
ev = maru_b200.Evaluator(model, gpu=0, max_batch=B, flags=maru_b200.capi.FLAG_NO_GRAPH)
dev = torch.device('cuda', 0)
d_states = torch.from_numpy(synth.random_states(B, 19, seed=7)).to(dev)
d_out = torch.empty((B, maru_b200.OUTPUT_SIZE), dtype=torch.float32, device=dev)
stream = torch.cuda.ExternalStream(ev.stream(), device=dev)
for bits in variants:
    ev.debug_set(1, bits)
    for _ in range(5): ev.forward_states_device(d_states.data_ptr(), d_out.data_ptr(), B)
    ev.sync()
    best = 1e9
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream): e0.record(stream)
        for _ in range(20): ev.forward_states_device(d_states.data_ptr(), d_out.data_ptr(), B)
        with torch.cuda.stream(stream): e1.record(stream)
        ev.sync()
        best = min(best, e0.elapsed_time(e1) / 20)
    import hashlib
    print(f'bits {bits:5d}: forward {best*1e3:7.1f} us   output sha {hashlib.sha1(d_out.cpu().numpy().tobytes()).hexdigest()[:12]}', flush=True)
ev.close()
