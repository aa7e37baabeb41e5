"""Per-launch times of one forward (maru_eval_profile_states: each launch repeated 10x inside one event pair).
usage: prof_layers.py [model] [batch] [debug bits]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import maru_b200
from bench import model_files
from maru_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else 'b4c128'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
bits = int(sys.argv[3]) if len(sys.argv) > 3 else 0
model, _ = model_files(name)
ev = maru_b200.Evaluator(model, gpu=0, max_batch=B)
ev.debug_set(1, bits)
st = torch.from_numpy(synth.random_states(B, 19, seed=7)).cuda()
out = torch.empty((B, maru_b200.OUTPUT_SIZE), dtype=torch.float32, device='cuda')
acc = {}
for rep in range(3):
    for i, l in enumerate(ev.profile_states_device(st.data_ptr(), out.data_ptr(), B)):
        if rep: acc.setdefault((i, l['name']), []).append((l['ms'], l['flops']))
tot = 0.0
for (i, n), v in acc.items():
    ms = sum(x[0] for x in v) / len(v); fl = v[0][1]; tot += ms
    if i < 14 or i > len(acc) - 4 or 'att0' in n: print(f'{i:3d} {n:28s} {ms*1e3:8.1f} us  {fl/ms/1e9 if fl else 0:7.1f} TF')
print(f'sum of launches {tot*1e3:.1f} us')
ev.close()
