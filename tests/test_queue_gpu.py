"""GPU: the leaf-batch queue (Processor/Executor mirror): many threads submitting n=1 jobs like
Evaluator.cpp:52 does, mixed job sizes and kinds, results identical to a direct forward."""
import threading

import numpy as np
import pytest

from conftest import load_positions

pytestmark = pytest.mark.gpu


def test_queue_matches_direct_and_batches(built_lib, model_dir, monkeypatch):
    import maru_b200
    # 19x19 and 9x9 positions travel in whatever batches the queue forms: with the compact frame for uniform batches
    # of small boards a 9x9 result would depend (in the last bits) on its batch mates, and this test compares bit for
    # bit — so every batch stays on the reference's 19x19 frame here (tests/test_compact_gpu.py covers the other one)
    monkeypatch.setenv('MARU_B200_COMPACT', '0')
    path = model_dir('b2c64')
    rows, states = load_positions('19')
    rows = np.concatenate([rows, load_positions('9')[0]])
    states = np.concatenate([states, load_positions('9')[1]])
    direct = maru_b200.Evaluator(path, gpu=0, max_batch=64)
    # two evaluators share the GPU below, which turns the persistent flow kernels off (their CTAs spin on each
    # other and need the device to themselves): compare bit for bit against the same one-launch-per-layer path
    direct.debug_set(1, 32)
    want = direct.forward(rows)
    direct.close()
    proc = maru_b200.Processor(path, gpus=[0], batch_size=32, threads_par_gpu=2)
    try:
        # Processor.execute with N rows (processor.py:34-39)
        assert np.array_equal(proc.execute(rows), want)
        assert np.array_equal(proc.execute_states(states), want)
        # 16 "search threads" each submitting single positions (Evaluator.cpp:52)
        got = np.zeros_like(want)
        def worker(tid):
            for i in range(tid, len(rows), 16):
                if i % 2:
                    got[i] = proc.execute(rows[i:i + 1])[0]
                else:
                    got[i] = proc.execute_states(states[i:i + 1])[0]
        ts = [threading.Thread(target=worker, args=(t,)) for t in range(16)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert np.array_equal(got, want)
        st = proc.queue.stats()
        assert st.positions >= 3 * len(rows) and st.batches >= 3
        # a job larger than the cap is never split across batches but still runs (Executor.cpp:116)
        big = proc.execute(np.concatenate([rows, rows]))
        assert np.array_equal(big[:len(rows)], want)
    finally:
        proc.close()


def test_queue_on_the_flow_kernels(built_lib, model_dir):
    """One evaluator per GPU (the default): the queue's batches of every size replay graphs built from the
    persistent conv_flow kernels (live batch size read on the device). Results must equal the direct call on the
    same path bit for bit, whatever batch a position happened to travel in."""
    import os
    import maru_b200
    path = model_dir('b4c128')
    rows, states = load_positions('19')
    os.environ['MARU_B200_FLOW'] = '1'        # the flow kernels at every batch size (default: for 249..354 positions on 148 SMs)
    try:
        direct = maru_b200.Evaluator(path, gpu=0, max_batch=64)
        want = direct.forward_states(states)
        assert direct.info().launches_per_forward <= 8            # encode, 3 flow launches, 2 attention, heads
        direct.close()
        proc = maru_b200.Processor(path, gpus=[0], batch_size=48, threads_par_gpu=1)
    finally:
        del os.environ['MARU_B200_FLOW']
    try:
        assert np.array_equal(proc.execute_states(states), want)
        got = np.zeros_like(want)
        def worker(tid):
            for i in range(tid, len(states), 12):
                k = 1 + (i % 3)                                # jobs of 1..3 positions
                got[i:i + k] = proc.execute_states(states[i:i + k])[:k]
        ts = [threading.Thread(target=worker, args=(t,)) for t in range(12)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert np.array_equal(got, want)
        assert proc.queue.stats().batches >= 3
    finally:
        proc.close()


def test_many_outstanding_device_forwards_with_varying_n(built_lib, model_dir):
    """More than 16 forwards of DIFFERENT live batch sizes enqueued on the stream without a sync: the live n of a
    graph replay travels by value (a ring of 16 pinned ints once let a later call overwrite an earlier one's n,
    which silently left rows uncomputed)."""
    import torch
    import maru_b200
    path = model_dir('b2c64')
    rows, states = load_positions('19')
    states = np.concatenate([states] * 4)[:8 * 22 + 3]          # 23 chunks of max_batch 8, the last one of 3
    ev = maru_b200.Evaluator(path, gpu=0, max_batch=8)
    try:
        want = ev.forward_states(states)                        # blocking, chunk by chunk
        d_states = torch.from_numpy(states).cuda()
        d_out = torch.zeros((len(states), maru_b200.OUTPUT_SIZE), dtype=torch.float32, device='cuda')
        torch.cuda.synchronize()
        # ragged sizes back to back, all asynchronous: 5, 8, 1, 7, ... (each call is its own graph replay)
        off = 0
        sizes = []
        while off < len(states):
            m = min([5, 8, 1, 7, 3, 8, 2][len(sizes) % 7], len(states) - off)
            ev.forward_states_device(d_states[off:].data_ptr(), d_out[off:].data_ptr(), m)
            sizes.append(m)
            off += m
        assert len(sizes) > 16
        ev.sync()
        assert np.array_equal(d_out.cpu().numpy(), want)
        d_out.zero_()
        ev.forward_states_device(d_states.data_ptr(), d_out.data_ptr(), len(states))   # one call, 23 chunks inside
        ev.sync()
        assert np.array_equal(d_out.cpu().numpy(), want)
    finally:
        ev.close()
