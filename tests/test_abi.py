"""CPU: the C-ABI library loads and exports every symbol include/maru_b200.h declares; creating an
evaluator without a GPU fails loudly (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / 'include' / 'maru_b200.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(maru_[a-z_0-9]+)\s*\(', text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ['maru_eval_create', 'maru_eval_forward_planes', 'maru_eval_forward_states',
              'maru_queue_execute', 'maru_encode_planes', 'maru_last_error']:
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert missing == []
    assert lib.maru_abi_version() == 1


def test_python_binding_lists_the_same_symbols(built_lib):
    from maru_b200.capi import ABI_SYMBOLS
    assert sorted(ABI_SYMBOLS) == declared_symbols()


def test_state_struct_is_448_bytes():
    text = (ROOT / 'include' / 'maru_b200.h').read_text()
    assert 'sizeof == 448' in text
    from maru_b200 import STATE_SIZE
    assert STATE_SIZE == 448 == 368 + 8 + 12 + 4 + 8 + 48


def test_no_cpu_fallback(built_lib, model_dir):
    import maru_b200
    from conftest import has_gpu
    with pytest.raises(maru_b200.MaruB200Error):
        maru_b200.Processor(model_dir('b2c64'), gpus=[-1])
    with pytest.raises(FileNotFoundError):
        maru_b200.Processor('/nonexistent/x.model', gpus=[0])
    if not has_gpu():
        with pytest.raises(maru_b200.MaruB200Error):
            maru_b200.Evaluator(model_dir('b2c64'), gpu=0, max_batch=8)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pat = re.compile(r'from\s+oracle|import\s+oracle|oracle/|liboracle|libmaru_ref|oracle\.')
    for p in (ROOT / 'maru_b200').rglob('*'):
        if p.suffix in ('.py', '.cu', '.cuh', '.cpp', '.h') and p.is_file():
            m = pat.search(p.read_text())
            assert m is None, f'{p} references the oracle: {m.group(0)}'
