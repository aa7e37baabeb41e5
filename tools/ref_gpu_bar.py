"""Same-box GPU bar: the UNMODIFIED reference runtime — deepgo::Processor -> Executor ->
Model::forward on libtorch CUDA, compiled from the reference sources into oracle/_ref — evaluating the same
TorchScript b4c128 file on the B200, fp32 and fp16, batch 256 of reference fp32 rows on the host
(the reference's own contract: 47 680 B per position H2D, 8 676 B D2H, synchronous copies)."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402,F401
import maru_b200  # noqa: E402
from bench import model_files  # noqa: E402
from maru_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model, _ = model_files('b4c128')
ev = maru_b200.Evaluator(model, gpu=0, max_batch=B)
states = synth.random_states(B, 19, seed=1234)
rows = ev.encode_planes(states)
ours = ev.forward(rows)
ev.close()
for fp16 in (False, True):
    proc = ref.RefProcessor(str(model), gpus=[0], batch_size=B, fp16=fp16, deterministic=False)
    out = proc.execute(rows)                      # warm-up (cudnn.benchmark tunes per batch size)
    for _ in range(5):
        proc.execute(rows)
    t0 = time.perf_counter()
    reps = 30
    for _ in range(reps):
        out = proc.execute(rows)
    dt = (time.perf_counter() - t0) / reps
    mask = rows[:, 32 * 361:33 * 361] > 0
    lp = lambda o: np.where(mask, np.log(np.maximum(o[:, :361], 1e-30)), 0.0)
    cl = lambda o: np.where(mask, lp(o) - (lp(o).sum(1) / mask.sum(1))[:, None], 0.0)
    print(f'reference on libtorch CUDA, fp16={fp16}: {B / dt:.0f} evals/s ({dt * 1e3:.3f} ms per batch of {B}); '
          f'vs maru_b200 outputs: centred log-policy max-abs {np.abs(cl(out) - cl(ours)).max():.2e}, '
          f'value max-abs {np.abs(out[:, 2166] - ours[:, 2166]).max():.2e}')
    proc.close()
