"""Short profiling target: a few b4c128 batch-256 forwards through the C ABI (used under ncu)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import maru_b200  # noqa: E402
from bench import model_files  # noqa: E402
from maru_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'b4c128'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 1   # 1 = no CUDA graph (plain launches)
model, _ = model_files(name)
ev = maru_b200.Evaluator(model, gpu=0, max_batch=B, flags=flags)
st = torch.from_numpy(synth.random_states(B, 19, seed=7)).cuda()
out = torch.empty((B, maru_b200.OUTPUT_SIZE), dtype=torch.float32, device='cuda')
for _ in range(reps):
    ev.forward_states_device(st.data_ptr(), out.data_ptr(), B)
ev.sync()
print('ok', float(out[:, :361].sum()))
ev.close()
