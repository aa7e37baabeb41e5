"""CPU: ThreadSanitizer run of the host-side concurrency of the hot path (leaf-batch queue: blocking jobs,
asynchronous tickets, coalescing; self-play scheduler workers) — `make -C maru_b200/csrc tsan` compiles
queue.cpp and selfplay.cpp unchanged with -fsanitize=thread over a host-only stub evaluator
(tests/tsan_queue_selfplay.cpp, maru_b200/csrc/engine_stub.h).  The reference has no race detection
(SURVEY 5); its queue protocol (Executor.cpp:52-138, ExecutorJob.cpp:8-36) is what the queue mirrors."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_queue_and_scheduler_are_race_free_under_tsan():
    if shutil.which('g++') is None:
        pytest.skip('no g++')
    r = subprocess.run(['make', '-C', str(ROOT / 'maru_b200' / 'csrc'), 'tsan'], capture_output=True, text=True)
    if r.returncode != 0 and 'sanitize' in (r.stderr + r.stdout):
        pytest.skip('this toolchain has no ThreadSanitizer runtime')
    assert r.returncode == 0, r.stderr[-2000:]
    run = subprocess.run([str(ROOT / 'maru_b200' / 'lib' / 'tsan_queue_selfplay')], capture_output=True, text=True,
                         timeout=300)
    assert 'ThreadSanitizer' not in run.stderr, run.stderr[-4000:]
    assert run.returncode == 0, (run.stdout, run.stderr[-2000:])
    assert '0 mismatches' in run.stdout
