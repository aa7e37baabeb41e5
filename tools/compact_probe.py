"""Bring-up probe: where does the compact-frame evaluation of a small board diverge from the 19x19-frame one?
Stops the launch sequence after k trunk launches (debug key 0) on both paths and compares the activation buffers."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))
import numpy as np  # noqa: E402
import maru_b200  # noqa: E402
from bench import model_files  # noqa: E402
from conftest import load_positions  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else '5x7'
name = sys.argv[2] if len(sys.argv) > 2 else 'b2c64'
model, _ = model_files(name)
rows, states = load_positions(tag)
states = states[:16]
n = len(states)
ev = maru_b200.Evaluator(model, gpu=0, max_batch=64)
C = ev.info().channels
bufs = {1: ('h', C), 2: ('r', C // 2), 3: ('s', C // 2), 4: ('qkv', 3 * C), 5: ('o', C), 6: ('g', 32)}
for k in range(1, 20):
    res = {}
    for mode in (0, 1 << 21):
        ev.debug_set(1, mode)
        ev.debug_set(0, k + (1 if mode == 0 else 0))      # the compact path has one more launch (crop) after the stem
        ev.forward_states(states)
        res[mode] = {b: ev.debug_planes(b, n, ch) for b, (_, ch) in bufs.items()}
    line = [f'after {k:2d} launches:']
    for b, (nm, _) in bufs.items():
        d = np.abs(res[0][b] - res[1 << 21][b]).max()
        line.append(f'{nm} {d:.3g}')
    print(' '.join(line), flush=True)
ev.close()
