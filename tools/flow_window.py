"""Forward time of the flow kernels against one launch per layer over a list of batch sizes (and board sizes): where does
the persistent path pay?  usage: flow_window.py [model] [board] batch [batch ...]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import maru_b200
from bench import model_files
from maru_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else 'b4c128'
board = int(sys.argv[2]) if len(sys.argv) > 2 else 19
batches = [int(a) for a in sys.argv[3:]] or [256]
model, _ = model_files(name)
dev = torch.device('cuda', 0)
for B in batches:
    ev = maru_b200.Evaluator(model, gpu=0, max_batch=B)
    if board != 19: ev.set_device_board(board, board)
    d_states = torch.from_numpy(synth.random_states(B, board, seed=7)).to(dev)
    d_out = torch.empty((B, maru_b200.OUTPUT_SIZE), dtype=torch.float32, device=dev)
    stream = torch.cuda.ExternalStream(ev.stream(), device=dev)
    res = []
    for bits in (128, 32):            # force the flow kernels / force one launch per layer
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
        res.append(best * 1e3)
    print(f'{name} board {board} batch {B:5d}: flow {res[0]:8.1f} us   per layer {res[1]:8.1f} us   ratio {res[0]/res[1]:.3f}', flush=True)
    ev.close()
