"""Per-step timeline of ONE CTA of the attention kernel (EXPERIMENTS build of the library, debug key 6):
softmax group (warp 0 / 4, lane 0): 0 = starts waiting for S, 1 = S ready, 2 = last S columns read out, 3 = P written + arrived;
control warp: 4 = P seen, 5 = PV issued, 6 = QK(step + 2) issued.   clock64 of the CTA's SM, printed relative to the first stamp.
  make -C maru_b200/csrc EXPERIMENTS=1 OUT=$PWD/maru_b200/lib_exp && MARU_B200_LIB=maru_b200/lib_exp/libmaru_b200.so python tools/att_trace.py [cta]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import maru_b200  # noqa: E402
from bench import model_files  # noqa: E402
from maru_b200 import synth  # noqa: E402

cta = int(sys.argv[1]) if len(sys.argv) > 1 else 500
B = 256
model, _ = model_files('b4c128')
ev = maru_b200.Evaluator(model, gpu=0, max_batch=B, flags=1)
st = torch.from_numpy(synth.random_states(B, 19, seed=7)).cuda()
out = torch.empty((B, maru_b200.OUTPUT_SIZE), dtype=torch.float32, device='cuda')
for _ in range(3):
    ev.forward_states_device(st.data_ptr(), out.data_ptr(), B)
ev.sync()
ev.debug_set(6, cta)
ev.forward_states_device(st.data_ptr(), out.data_ptr(), B)
ev.sync()
t = ev.debug_trace(4).reshape(-1)[:64 * 16].reshape(64, 16)
t0 = t[t > 0].min()
names = ['wait S', 'S ready', 'S read', 'P done', 'ctl: P seen', 'PV issued', 'QK issued']
print('step ' + ' '.join(f'{n:>11s}' for n in names))
for s in range(16):
    if t[s].max() == 0:
        continue
    print(f'{s:4d} ' + ' '.join(f'{(int(v - t0) if v > 0 else -1):11d}' for v in t[s, :7]))
print('epilogue of tile 0..2 (group 0, warp 0): start, O ready, O read, stored:')
for k in range(3):
    print('   ', [int(v - t0) if v > 0 else -1 for v in t[32 + k, :4]])
d = np.diff(np.sort(t[:15, 3][t[:15, 3] > 0]))
print('P-done to P-done (cycles):', d.tolist(), 'mean', float(d.mean()) if len(d) else None)
ev.close()
