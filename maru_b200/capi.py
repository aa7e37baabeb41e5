"""ctypes bindings of include/maru_b200.h and the host-side mirror of the reference's Python API.

`Processor` mirrors reference src/deepgo/processor.py:9-39 (same constructor arguments, same
`execute(np.float32[N,11920]) -> np.float32[N,2169]`, same FileNotFoundError), so code written
against `deepgo.processor.Processor` runs unchanged; `execute_states` is the added compact path.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

INPUT_SIZE = 11920    # reference src/deepgo/config.py:48
OUTPUT_SIZE = 2169    # reference src/deepgo/config.py:50
STATE_SIZE = 448      # sizeof(maru_state)
LEAF_SIZE = 364       # sizeof(maru_leaf) / 4: prior[361], value, score[2]
FLAG_NO_GRAPH = 1


class MaruB200Error(RuntimeError):
    pass


class EvalInfo(C.Structure):
    _fields_ = [('blocks', C.c_int32), ('channels', C.c_int32), ('attn_every', C.c_int32),
                ('heads', C.c_int32), ('head_channels', C.c_int32), ('value_channels', C.c_int32),
                ('n_attn', C.c_int32), ('gpu_id', C.c_int32), ('max_batch', C.c_int32),
                ('sm_count', C.c_int32), ('launches_per_forward', C.c_int32),
                ('reserved0', C.c_int32), ('flops_per_eval', C.c_double),
                ('flops_per_eval_executed', C.c_double), ('n_forward', C.c_uint64),
                ('n_positions', C.c_uint64), ('device_ms_last', C.c_double)]


class LaunchTime(C.Structure):
    _fields_ = [('name', C.c_char * 32), ('kind', C.c_int32), ('cin', C.c_int32),
                ('cout', C.c_int32), ('taps', C.c_int32), ('ms', C.c_float),
                ('flops', C.c_double), ('bytes', C.c_double)]


class QueueStats(C.Structure):
    _fields_ = [('jobs', C.c_uint64), ('positions', C.c_uint64), ('batches', C.c_uint64),
                ('max_batch_seen', C.c_uint64), ('wait_ms_total', C.c_double),
                ('device_ms_total', C.c_double)]


class SelfPlayConfig(C.Structure):
    """maru_selfplay_config (include/maru_b200.h)."""
    _fields_ = [('width', C.c_int32), ('height', C.c_int32), ('komi', C.c_float), ('rule', C.c_int32),
                ('superko', C.c_int32), ('games', C.c_int32), ('threads', C.c_int32), ('visits', C.c_int32),
                ('max_moves', C.c_int32), ('root', C.c_int32), ('root_width', C.c_int32), ('noise', C.c_float),
                ('temperature', C.c_float), ('criterion', C.c_int32), ('eval_leaf_only', C.c_int32),
                ('opening_moves', C.c_int32), ('restart', C.c_int32), ('seed', C.c_uint64),
                ('cpu_first', C.c_int32), ('cpu_count', C.c_int32), ('lanes', C.c_int32), ('record', C.c_int32),
                ('use_ucb1', C.c_int32)]


class SelfPlayStats(C.Structure):
    _fields_ = [('visits', C.c_uint64), ('evals', C.c_uint64), ('expansions', C.c_uint64), ('moves', C.c_uint64),
                ('games_finished', C.c_uint64), ('submits', C.c_uint64), ('max_submit', C.c_uint64),
                ('run_s', C.c_double), ('host_s', C.c_double), ('stall_s', C.c_double), ('device_s', C.c_double)]


class CandidateRec(C.Structure):
    _fields_ = [('x', C.c_int32), ('y', C.c_int32), ('color', C.c_int32), ('visits', C.c_int32),
                ('playouts', C.c_int32), ('prior', C.c_float), ('value', C.c_float)]


LEAF_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)   # maru_leaf_fn


# every symbol include/maru_b200.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    'maru_eval_create', 'maru_eval_destroy', 'maru_eval_get_info', 'maru_eval_forward_planes',
    'maru_eval_forward_states', 'maru_eval_forward_states_device', 'maru_eval_forward_leaves',
    'maru_eval_forward_leaves_device', 'maru_queue_execute_leaves',
    'maru_eval_forward_planes_device', 'maru_eval_sync', 'maru_eval_stream', 'maru_encode_planes',
    'maru_encode_planes_device', 'maru_eval_debug_planes', 'maru_eval_debug_policy_logits',
    'maru_eval_debug_set', 'maru_eval_debug_trace', 'maru_eval_debug_timeline', 'maru_eval_debug_stamps', 'maru_eval_profile_states',
    'maru_queue_create', 'maru_queue_destroy', 'maru_queue_execute', 'maru_queue_execute_states',
    'maru_queue_get_stats', 'maru_ladder_flags', 'maru_rules_eval', 'maru_game_create', 'maru_game_clone',
    'maru_game_destroy', 'maru_game_play', 'maru_game_state', 'maru_last_error', 'maru_abi_version',
    'maru_queue_submit_leaves', 'maru_queue_poll', 'maru_queue_wait', 'maru_queue_set_coalesce',
    'maru_eval_set_device_board', 'maru_eval_lane_create', 'maru_lane_destroy', 'maru_lane_states', 'maru_lane_leaves',
    'maru_lane_submit', 'maru_lane_poll', 'maru_lane_wait', 'maru_selfplay_create_eval',
    'maru_selfplay_create', 'maru_selfplay_create_fn', 'maru_selfplay_destroy', 'maru_selfplay_play',
    'maru_selfplay_run', 'maru_selfplay_get_stats', 'maru_selfplay_get_moves', 'maru_selfplay_get_candidates',
]


def lib_path() -> Path:
    env = os.environ.get('MARU_B200_LIB')
    return Path(env) if env else Path(__file__).resolve().parent / 'lib' / 'libmaru_b200.so'


_lib = None


def load_library():
    """Load libmaru_b200.so (in-tree build: `make -C maru_b200/csrc`). Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not p.exists():
        raise MaruB200Error(f'{p} not found: build it with `make -C maru_b200/csrc` '
                            '(or python -c "import __graft_entry__ as g; g.build()")')
    lib = C.CDLL(str(p), mode=C.RTLD_GLOBAL)
    vp, ci, cu = C.c_void_p, C.c_int, C.c_uint
    lib.maru_eval_create.argtypes = [C.c_char_p, ci, ci, cu, C.POINTER(vp)]
    lib.maru_eval_destroy.argtypes = [vp]
    lib.maru_eval_destroy.restype = None
    lib.maru_eval_get_info.argtypes = [vp, C.POINTER(EvalInfo)]
    lib.maru_eval_forward_planes.argtypes = [vp, vp, vp, ci]
    lib.maru_eval_forward_states.argtypes = [vp, vp, vp, ci]
    lib.maru_eval_forward_states_device.argtypes = [vp, vp, vp, ci]
    lib.maru_eval_forward_leaves.argtypes = [vp, vp, vp, ci]
    lib.maru_eval_forward_leaves_device.argtypes = [vp, vp, vp, ci]
    lib.maru_queue_execute_leaves.argtypes = [vp, vp, vp, ci]
    lib.maru_eval_forward_planes_device.argtypes = [vp, vp, vp, ci]
    lib.maru_eval_sync.argtypes = [vp]
    lib.maru_eval_set_device_board.argtypes = [vp, ci, ci]
    lib.maru_eval_stream.argtypes = [vp]
    lib.maru_eval_stream.restype = vp
    lib.maru_encode_planes.argtypes = [vp, vp, vp, ci]
    lib.maru_encode_planes_device.argtypes = [vp, vp, vp, ci]
    lib.maru_eval_debug_planes.argtypes = [vp, ci, vp, ci, ci]
    lib.maru_eval_debug_set.argtypes = [vp, ci, ci]
    lib.maru_eval_debug_trace.argtypes = [vp, vp, ci]
    lib.maru_eval_debug_timeline.argtypes = [vp, vp]
    lib.maru_eval_debug_stamps.argtypes = [vp, vp]
    lib.maru_eval_profile_states.argtypes = [vp, vp, vp, ci, C.POINTER(LaunchTime), ci,
                                             C.POINTER(ci)]
    lib.maru_eval_debug_policy_logits.argtypes = [vp, vp, ci]
    lib.maru_queue_create.argtypes = [C.POINTER(vp), ci, ci, C.POINTER(vp)]
    lib.maru_queue_destroy.argtypes = [vp]
    lib.maru_queue_destroy.restype = None
    lib.maru_queue_execute.argtypes = [vp, vp, vp, ci]
    lib.maru_queue_execute_states.argtypes = [vp, vp, vp, ci]
    lib.maru_queue_get_stats.argtypes = [vp, C.POINTER(QueueStats)]
    lib.maru_ladder_flags.argtypes = [vp, ci, ci, ci, ci, ci, vp]
    lib.maru_rules_eval.argtypes = [vp, ci, ci, ci, ci, ci, ci, vp, vp]
    lib.maru_game_create.argtypes = [ci, ci]
    lib.maru_game_create.restype = vp
    lib.maru_game_clone.argtypes = [vp]
    lib.maru_game_clone.restype = vp
    lib.maru_game_destroy.argtypes = [vp]
    lib.maru_game_destroy.restype = None
    lib.maru_game_play.argtypes = [vp, ci, ci, ci]
    lib.maru_game_state.argtypes = [vp, ci, C.c_float, ci, ci, ci, vp]
    lib.maru_queue_submit_leaves.argtypes = [vp, vp, vp, ci, C.POINTER(vp)]
    lib.maru_queue_poll.argtypes = [vp, vp, C.POINTER(ci)]
    lib.maru_queue_wait.argtypes = [vp, vp]
    lib.maru_queue_set_coalesce.argtypes = [vp, ci, ci]
    lib.maru_selfplay_create.argtypes = [vp, C.POINTER(SelfPlayConfig), C.POINTER(vp)]
    lib.maru_selfplay_create_fn.argtypes = [vp, vp, C.POINTER(SelfPlayConfig), C.POINTER(vp)]
    lib.maru_selfplay_create_eval.argtypes = [vp, C.POINTER(SelfPlayConfig), C.POINTER(vp)]
    lib.maru_eval_lane_create.argtypes = [vp, ci, C.POINTER(vp)]
    lib.maru_lane_destroy.argtypes = [vp]
    lib.maru_lane_destroy.restype = None
    lib.maru_lane_states.argtypes = [vp]
    lib.maru_lane_states.restype = vp
    lib.maru_lane_leaves.argtypes = [vp]
    lib.maru_lane_leaves.restype = vp
    lib.maru_lane_submit.argtypes = [vp, ci]
    lib.maru_lane_poll.argtypes = [vp, C.POINTER(ci)]
    lib.maru_lane_wait.argtypes = [vp]
    lib.maru_selfplay_destroy.argtypes = [vp]
    lib.maru_selfplay_destroy.restype = None
    lib.maru_selfplay_play.argtypes = [vp, ci, ci, ci]
    lib.maru_selfplay_run.argtypes = [vp, ci]
    lib.maru_selfplay_get_stats.argtypes = [vp, C.POINTER(SelfPlayStats)]
    lib.maru_selfplay_get_moves.argtypes = [vp, ci, vp, ci]
    lib.maru_selfplay_get_candidates.argtypes = [vp, ci, ci, C.POINTER(CandidateRec), ci]
    lib.maru_last_error.restype = C.c_char_p
    lib.maru_abi_version.restype = ci
    _lib = lib
    return lib


def _check(lib, rc: int, what: str) -> None:
    if rc != 0:
        raise MaruB200Error(f'{what}: {lib.maru_last_error().decode(errors="replace")}')


def ladder_flags(colors: np.ndarray, ko: Optional[tuple] = None) -> np.ndarray:
    """Exact Board::isShicho flags of a position (host code). colors: [h, w] int8, 0 / +1 black / -1 white;
    ko = (x, y, color) if `color` may not retake at (x, y). Returns uint8 [h, w]."""
    lib = load_library()
    col = np.ascontiguousarray(colors, np.int8)
    h, w = col.shape
    out = np.zeros((h, w), np.uint8)
    kx, ky, kc = ko if ko is not None else (-1, -1, 0)
    _check(lib, lib.maru_ladder_flags(col.ctypes.data, w, h, kx, ky, kc, out.ctypes.data), 'maru_ladder_flags')
    return out


def rules_eval(colors: np.ndarray, color: int, ko: Optional[tuple] = None):
    """(enableds uint8 [h, w] = Board::getEnableds(color, checkSeki=True), territories int8 [h, w] =
    Board::getTerritories(color)) of a position, on the flat board (host code)."""
    lib = load_library()
    col = np.ascontiguousarray(colors, np.int8)
    h, w = col.shape
    en = np.zeros((h, w), np.uint8)
    te = np.zeros((h, w), np.int8)
    kx, ky, kc = ko if ko is not None else (-1, -1, 0)
    _check(lib, lib.maru_rules_eval(col.ctypes.data, w, h, kx, ky, kc, color, en.ctypes.data, te.ctypes.data),
           'maru_rules_eval')
    return en, te


class LeanGame:
    """A position on the flat host board (maru_game_*): Board::play semantics and the compact record,
    byte-identical to the reference Board + state_from_board, without building a reference Board."""

    def __init__(self, width: int = 19, height: int = 19, _handle=None):
        self.lib = load_library()
        self.width, self.height = width, height
        self.handle = _handle if _handle is not None else self.lib.maru_game_create(width, height)
        if not self.handle:
            raise MaruB200Error(self.lib.maru_last_error().decode(errors='replace') or 'maru_game_create failed')

    def copy(self) -> 'LeanGame':
        return LeanGame(self.width, self.height, self.lib.maru_game_clone(self.handle))

    def play(self, x: int, y: int, color: int) -> int:
        return self.lib.maru_game_play(self.handle, x, y, color)

    def state(self, color: int, komi: float = 7.5, rule: int = 0, superko: bool = False,
              with_mask: bool = False) -> np.ndarray:
        out = np.zeros(STATE_SIZE, np.uint8)
        _check(self.lib, self.lib.maru_game_state(self.handle, color, float(komi), rule, int(superko),
                                                  int(with_mask), out.ctypes.data), 'maru_game_state')
        return out

    def close(self) -> None:
        if getattr(self, 'handle', None):
            self.lib.maru_game_destroy(self.handle)
            self.handle = None

    def __del__(self):
        self.close()


def ensure_pack(model: str | Path) -> Path:
    """Return the `.b200` weight pack of a `.model` file, converting it if it is missing or stale."""
    model = Path(model)
    if model.suffix == '.b200':
        return model
    pack = Path(str(model) + '.b200')
    if not pack.exists() or pack.stat().st_mtime < model.stat().st_mtime:
        from .convert import convert  # the loader is the only torch user
        convert(str(model), str(pack))
    return pack


def _rows(x: np.ndarray, width: int, dtype) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=dtype)
    if x.ndim == 1:
        x = x.reshape(1, -1)
    if x.ndim != 2 or x.shape[1] != width:
        raise ValueError(f'expected [N,{width}] {np.dtype(dtype).name}, got {x.shape}')
    return x


class Evaluator:
    """One evaluator = one GPU, one stream, one weight replica (mirrors deepgo::Model,
    reference src/deepgo/native/cpp/Model.h:12-58)."""

    def __init__(self, model: str | Path, gpu: int = 0, max_batch: int = 256, flags: int = 0):
        if not Path(model).exists():
            raise FileNotFoundError(f'File not found: {model}')
        self.lib = load_library()
        pack = ensure_pack(model)
        h = C.c_void_p()
        _check(self.lib, self.lib.maru_eval_create(str(pack).encode(), gpu, max_batch, flags,
                                                   C.byref(h)), 'maru_eval_create')
        self.handle = h
        self.max_batch = max_batch

    def close(self) -> None:
        if getattr(self, 'handle', None):
            self.lib.maru_eval_destroy(self.handle)
            self.handle = None

    def __del__(self):
        self.close()

    def info(self) -> EvalInfo:
        inf = EvalInfo()
        self.lib.maru_eval_get_info(self.handle, C.byref(inf))
        return inf

    def forward(self, inputs: np.ndarray) -> np.ndarray:
        """[N,11920] fp32 reference rows -> [N,2169] fp32 (Model::forward, Model.cpp:88-100)."""
        x = _rows(inputs, INPUT_SIZE, np.float32)
        out = np.zeros((x.shape[0], OUTPUT_SIZE), np.float32)
        _check(self.lib, self.lib.maru_eval_forward_planes(self.handle, x.ctypes.data,
                                                           out.ctypes.data, x.shape[0]),
               'maru_eval_forward_planes')
        return out

    def forward_states(self, states: np.ndarray) -> np.ndarray:
        """[N,448] uint8 compact records -> [N,2169] fp32; encoding runs on the GPU (K1)."""
        s = _rows(states, STATE_SIZE, np.uint8)
        out = np.zeros((s.shape[0], OUTPUT_SIZE), np.float32)
        _check(self.lib, self.lib.maru_eval_forward_states(self.handle, s.ctypes.data,
                                                           out.ctypes.data, s.shape[0]),
               'maru_eval_forward_states')
        return out

    def forward_leaves(self, states: np.ndarray) -> np.ndarray:
        """[N,448] compact records -> [N,364] fp32 compact results (maru_leaf: prior[361] in board
        order y*width+x, masked by the record's move_mask; value; score[2]) — what
        Evaluator::evaluate keeps of the output row (Evaluator.cpp:55-80)."""
        s = _rows(states, STATE_SIZE, np.uint8)
        out = np.zeros((s.shape[0], LEAF_SIZE), np.float32)
        _check(self.lib, self.lib.maru_eval_forward_leaves(self.handle, s.ctypes.data,
                                                           out.ctypes.data, s.shape[0]),
               'maru_eval_forward_leaves')
        return out

    def encode_planes(self, states: np.ndarray) -> np.ndarray:
        """[N,448] compact records -> [N,11920] fp32 rows, bit-exact Board::getInputs."""
        s = _rows(states, STATE_SIZE, np.uint8)
        out = np.zeros((s.shape[0], INPUT_SIZE), np.float32)
        _check(self.lib, self.lib.maru_encode_planes(self.handle, s.ctypes.data, out.ctypes.data,
                                                     s.shape[0]), 'maru_encode_planes')
        return out

    def debug_planes(self, which: int, n: int, channels: int) -> np.ndarray:
        """Activation buffer `which` (MARU_BUF_*) of the last forward as fp32 [n][channels][361]."""
        out = np.zeros((n, channels, 361), np.float32)
        _check(self.lib, self.lib.maru_eval_debug_planes(self.handle, which, out.ctypes.data, n,
                                                         channels), 'maru_eval_debug_planes')
        return out

    def debug_trace(self, n_ctas: int = 148) -> np.ndarray:
        out = np.zeros((n_ctas, 32, 8), np.int64)
        _check(self.lib, self.lib.maru_eval_debug_trace(self.handle, out.ctypes.data, n_ctas),
               'maru_eval_debug_trace')
        return out

    def debug_timeline(self) -> np.ndarray:
        out = np.zeros((64, 2, 8), np.int64)
        _check(self.lib, self.lib.maru_eval_debug_timeline(self.handle, out.ctypes.data),
               'maru_eval_debug_timeline')
        return out

    def debug_stamps(self) -> np.ndarray:
        out = np.zeros((128, 4), np.int64)
        _check(self.lib, self.lib.maru_eval_debug_stamps(self.handle, out.ctypes.data), 'maru_eval_debug_stamps')
        return out

    def debug_set(self, key: int, value: int) -> None:
        _check(self.lib, self.lib.maru_eval_debug_set(self.handle, key, value),
               'maru_eval_debug_set')

    def debug_policy_logits(self, n: int) -> np.ndarray:
        out = np.zeros((n, 2, 361), np.float32)
        _check(self.lib, self.lib.maru_eval_debug_policy_logits(self.handle, out.ctypes.data, n),
               'debug_policy_logits')
        return out

    # device-resident entry points (raw device pointers as ints, e.g. torch_tensor.data_ptr())
    def forward_states_device(self, d_states: int, d_out: int, n: int) -> None:
        _check(self.lib, self.lib.maru_eval_forward_states_device(self.handle, d_states, d_out, n),
               'maru_eval_forward_states_device')

    def forward_leaves_device(self, d_states: int, d_out: int, n: int) -> None:
        _check(self.lib, self.lib.maru_eval_forward_leaves_device(self.handle, d_states, d_out, n),
               'maru_eval_forward_leaves_device')

    def forward_planes_device(self, d_in: int, d_out: int, n: int) -> None:
        _check(self.lib, self.lib.maru_eval_forward_planes_device(self.handle, d_in, d_out, n),
               'maru_eval_forward_planes_device')

    def encode_planes_device(self, d_states: int, d_rows: int, n: int) -> None:
        _check(self.lib, self.lib.maru_encode_planes_device(self.handle, d_states, d_rows, n),
               'maru_encode_planes_device')

    def profile_states_device(self, d_states: int, d_out: int, n: int):
        """One un-graphed forward with a CUDA-event pair around every launch -> list of dicts."""
        arr = (LaunchTime * 256)()
        cnt = C.c_int(0)
        _check(self.lib, self.lib.maru_eval_profile_states(self.handle, d_states, d_out, n, arr, 256,
                                                           C.byref(cnt)), 'maru_eval_profile_states')
        return [dict(name=arr[i].name.decode(), kind=arr[i].kind, cin=arr[i].cin, cout=arr[i].cout,
                     taps=arr[i].taps, ms=arr[i].ms, flops=arr[i].flops, bytes=arr[i].bytes)
                for i in range(min(cnt.value, 256))]

    def set_device_board(self, width: int, height: int) -> None:
        """Board size of the records the *_device entry points will be given (0, 0: unknown -> 19x19 frame)."""
        _check(self.lib, self.lib.maru_eval_set_device_board(self.handle, width, height),
               'maru_eval_set_device_board')

    def lane(self, max_rows: int) -> 'EvalLane':
        """Pinned staging of one in-flight batch (maru_lane_*): fill .states[:n], submit(n), wait() -> .leaves[:n]."""
        return EvalLane(self, max_rows)

    def sync(self) -> None:
        _check(self.lib, self.lib.maru_eval_sync(self.handle), 'maru_eval_sync')

    def stream(self) -> int:
        return int(self.lib.maru_eval_stream(self.handle) or 0)


class EvalLane:
    """maru_lane: the caller stages records in pinned memory and submits the batch itself (no queue thread)."""

    def __init__(self, ev: Evaluator, max_rows: int):
        self.lib, self.ev = ev.lib, ev
        h = C.c_void_p()
        _check(self.lib, self.lib.maru_eval_lane_create(ev.handle, max_rows, C.byref(h)), 'maru_eval_lane_create')
        self.handle = h
        self.states = np.ctypeslib.as_array(C.cast(self.lib.maru_lane_states(h), C.POINTER(C.c_uint8)),
                                            shape=(max_rows, STATE_SIZE))
        self.leaves = np.ctypeslib.as_array(C.cast(self.lib.maru_lane_leaves(h), C.POINTER(C.c_float)),
                                            shape=(max_rows, LEAF_SIZE))

    def submit(self, n: int) -> None:
        _check(self.lib, self.lib.maru_lane_submit(self.handle, n), 'maru_lane_submit')

    def poll(self) -> bool:
        d = C.c_int(0)
        _check(self.lib, self.lib.maru_lane_poll(self.handle, C.byref(d)), 'maru_lane_poll')
        return bool(d.value)

    def wait(self) -> None:
        _check(self.lib, self.lib.maru_lane_wait(self.handle), 'maru_lane_wait')

    def close(self) -> None:
        if getattr(self, 'handle', None):
            self.states = self.leaves = None
            self.lib.maru_lane_destroy(self.handle)
            self.handle = None


class LeafQueue:
    """Leaf-batch queue over one or more evaluators (maru_queue_*; mirrors deepgo::Processor +
    Executor, reference Processor.cpp:17-60 / Executor.cpp:52-189). Thread-safe, blocking."""

    def __init__(self, evaluators: Sequence[Evaluator], batch_size: int = 2048):
        self.lib = load_library()
        self.evaluators = list(evaluators)
        arr = (C.c_void_p * len(self.evaluators))(*[e.handle for e in self.evaluators])
        h = C.c_void_p()
        _check(self.lib, self.lib.maru_queue_create(arr, len(self.evaluators), batch_size,
                                                    C.byref(h)), 'maru_queue_create')
        self.handle = h

    def close(self) -> None:
        if getattr(self, 'handle', None):
            self.lib.maru_queue_destroy(self.handle)
            self.handle = None

    def __del__(self):
        self.close()

    def execute(self, inputs: np.ndarray) -> np.ndarray:
        x = _rows(inputs, INPUT_SIZE, np.float32)
        out = np.zeros((x.shape[0], OUTPUT_SIZE), np.float32)
        _check(self.lib, self.lib.maru_queue_execute(self.handle, x.ctypes.data, out.ctypes.data,
                                                     x.shape[0]), 'maru_queue_execute')
        return out

    def execute_states(self, states: np.ndarray) -> np.ndarray:
        s = _rows(states, STATE_SIZE, np.uint8)
        out = np.zeros((s.shape[0], OUTPUT_SIZE), np.float32)
        _check(self.lib, self.lib.maru_queue_execute_states(self.handle, s.ctypes.data,
                                                            out.ctypes.data, s.shape[0]),
               'maru_queue_execute_states')
        return out

    def execute_leaves(self, states: np.ndarray) -> np.ndarray:
        s = _rows(states, STATE_SIZE, np.uint8)
        out = np.zeros((s.shape[0], LEAF_SIZE), np.float32)
        _check(self.lib, self.lib.maru_queue_execute_leaves(self.handle, s.ctypes.data,
                                                            out.ctypes.data, s.shape[0]),
               'maru_queue_execute_leaves')
        return out

    def stats(self) -> QueueStats:
        st = QueueStats()
        self.lib.maru_queue_get_stats(self.handle, C.byref(st))
        return st

    # asynchronous variant: submit a batch of leaves, keep working, collect later
    def submit_leaves(self, states: np.ndarray):
        """-> ticket (opaque); the result array is returned by wait(ticket)."""
        s = _rows(states, STATE_SIZE, np.uint8)
        out = np.zeros((s.shape[0], LEAF_SIZE), np.float32)
        t = C.c_void_p()
        _check(self.lib, self.lib.maru_queue_submit_leaves(self.handle, s.ctypes.data, out.ctypes.data,
                                                           s.shape[0], C.byref(t)), 'maru_queue_submit_leaves')
        return (t, s, out)                       # the buffers must outlive the ticket

    def poll(self, ticket) -> bool:
        done = C.c_int(0)
        _check(self.lib, self.lib.maru_queue_poll(self.handle, ticket[0], C.byref(done)), 'maru_queue_poll')
        return bool(done.value)

    def wait(self, ticket) -> np.ndarray:
        _check(self.lib, self.lib.maru_queue_wait(self.handle, ticket[0]), 'maru_queue_wait')
        return ticket[2]

    def set_coalesce(self, min_rows: int, max_wait_us: int) -> None:
        _check(self.lib, self.lib.maru_queue_set_coalesce(self.handle, min_rows, max_wait_us),
               'maru_queue_set_coalesce')


class Processor:
    """Drop-in for `deepgo.processor.Processor` (reference src/deepgo/processor.py:9-39).

    Same arguments and behaviour: `gpus` lists CUDA ordinals (one evaluator per gpu x
    threads_par_gpu, least-loaded dispatch). Differences forced by the B200-only design:
    `gpus=[-1]` (CPU) raises instead of silently computing on the host; `fp16` is accepted and
    ignored (the trunk always runs bf16 tensor-core math with fp32 accumulation);
    `deterministic` is accepted (the kernels are run-to-run deterministic anyway)."""

    def __init__(self, model: str | Path, gpus: List[int] = [0], batch_size: int = 2048,
                 fp16: bool = False, deterministic: bool = False, threads_par_gpu: int = 1,
                 max_batch: Optional[int] = None) -> None:
        if not Path(model).exists():
            raise FileNotFoundError(f'File not found: {model}')
        if any(g < 0 for g in gpus):
            raise MaruB200Error('maru_b200 has no CPU path: gpus must be CUDA ordinals >= 0')
        mb = max_batch or min(batch_size, 2048)
        self.evaluators = [Evaluator(model, g, mb) for g in gpus for _ in range(threads_par_gpu)]
        self.queue = LeafQueue(self.evaluators, batch_size)

    def execute(self, inputs: np.ndarray) -> np.ndarray:
        return self.queue.execute(inputs)

    def execute_states(self, states: np.ndarray) -> np.ndarray:
        return self.queue.execute_states(states)

    def execute_leaves(self, states: np.ndarray) -> np.ndarray:
        return self.queue.execute_leaves(states)

    def close(self) -> None:
        self.queue.close()
        for e in self.evaluators:
            e.close()
