"""Run-to-run determinism probe: which trunk launch first produces different bits on a repeat."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import numpy as np
import maru_b200
from bench import model_files
from test_fullsize_gpu import mixed_states
name = sys.argv[1] if len(sys.argv) > 1 else 'b10c256'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
model, _ = model_files(name)
ev = maru_b200.Evaluator(model, gpu=0, max_batch=B, flags=flags)
st = mixed_states(B, seed=11)
outs = [ev.forward_states(st) for _ in range(4)]
for i in range(1, 4):
    d = np.argwhere((outs[i] != outs[0]).any(1)).ravel()
    print(f'run {i} vs 0: {len(d)} positions differ', d[:12].tolist())
info = ev.info()
C = info.channels
chan = {1: C, 2: C // 2, 3: C // 2, 4: 3 * C, 5: C, 6: info.head_channels}
# launch order: stem | per block: reduce c1 c2 c1 c2 expand | per attention: qkv core proj | head.g
seq = [('stem', 1)]
ai = 0
for b in range(info.blocks):
    seq += [(f'blk{b}.reduce', 2), (f'blk{b}.c1a', 3), (f'blk{b}.c2a', 2), (f'blk{b}.c1b', 3), (f'blk{b}.c2b', 2), (f'blk{b}.expand', 1)]
    if info.attn_every > 0 and (b + 1) % info.attn_every == 0 and ai < info.n_attn:
        seq += [(f'att{ai}.qkv', 4), (f'att{ai}.core', 5), (f'att{ai}.proj', 1)]
        ai += 1
seq += [('head.g', 6)]
for i, (nm, buf) in enumerate(seq):
    ev.debug_set(0, i + 1)
    got = []
    for _ in range(3):
        ev.forward_states(st)
        got.append(ev.debug_planes(buf, B, chan[buf]))
    bad = [int((got[k] != got[0]).sum()) for k in (1, 2)]
    nanc = int(np.isnan(got[0]).sum())
    if any(bad) or nanc:
        pos = np.argwhere((got[1] != got[0]).any((1, 2))).ravel()
        print(f'{i:3d} {nm:14s} differing elements {bad} nan {nanc} positions {pos[:10].tolist()}')
        if '-v' not in sys.argv:
            break
    else:
        print(f'{i:3d} {nm:14s} repeatable')
ev.debug_set(0, -1)
ev.close()
