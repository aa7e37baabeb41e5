import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / 'tests' / 'golden'
GOLDEN_TAGS = ['9', '13', '19', '9x13', '5x7']


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu on the GPU box)')


def load_positions(tag: str):
    """golden fixture -> (rows [n,11920] fp32 as the reference wrote them, states [n,448] uint8)."""
    z = np.load(GOLDEN / f'positions_{tag}.npz')
    planes = np.unpackbits(z['plane_bits'], axis=1)[:, :33 * 361].astype(np.float32)
    rows = np.concatenate([planes, z['scalars'].astype(np.float32)], axis=1)
    return np.ascontiguousarray(rows), np.ascontiguousarray(z['states'])


@pytest.fixture(scope='session')
def built_lib():
    """libmaru_b200.so, built in-tree if missing."""
    lib = ROOT / 'maru_b200' / 'lib' / 'libmaru_b200.so'
    if not lib.exists():
        subprocess.check_call(['make', '-C', str(ROOT / 'maru_b200' / 'csrc'), '-j8'])
    return lib


@pytest.fixture(scope='session')
def oracle_features_lib():
    import ctypes
    so = ROOT / 'oracle' / '_build' / 'liboracle_features.so'
    if not so.exists():
        subprocess.check_call(['make', '-C', str(ROOT / 'oracle'), 'features'])
    return ctypes.CDLL(str(so))


@pytest.fixture(scope='session')
def model_dir(tmp_path_factory):
    """Exports netspec random-init models (TorchScript .model + .b200 pack) once per session."""
    from maru_b200 import convert, netspec
    d = tmp_path_factory.mktemp('models')
    paths = {}

    def get(name: str, seed: int = 0):
        key = (name, seed)
        if key not in paths:
            p = d / f'{name}-seed{seed}.model'
            netspec.export(netspec.build(**netspec.CONFIGS[name], seed=seed), str(p))
            convert.convert(str(p))
            paths[key] = p
        return paths[key]
    return get


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
