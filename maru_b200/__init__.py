"""maru_b200 — B200-native leaf evaluator for takedarts/maru (one hot path, drop-in).

Python face of libmaru_b200.so (C ABI in include/maru_b200.h). PyTorch is imported only by
`maru_b200.convert` (the loader) and `maru_b200.netspec` (architecture spec / export tool);
evaluation itself is hand-written sm_100a CUDA behind ctypes. There is no CPU fallback: if the
shared library is missing or no B200 is visible, construction raises.
"""
from .capi import (INPUT_SIZE, LEAF_SIZE, OUTPUT_SIZE, STATE_SIZE, Evaluator, LeafQueue, MaruB200Error,
                   Processor, lib_path, load_library)

__all__ = ['INPUT_SIZE', 'LEAF_SIZE', 'OUTPUT_SIZE', 'STATE_SIZE', 'Evaluator', 'LeafQueue', 'MaruB200Error',
           'Processor', 'lib_path', 'load_library']
