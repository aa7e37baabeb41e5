"""TEST INFRASTRUCTURE ONLY (oracle). ctypes bindings to the UNMODIFIED reference, compiled by
oracle/Makefile into oracle/_ref/libmaru_ref_{board,nn,dropin}.so (see oracle/ref_shim.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module. The product (maru_b200/) never does.

Reference entry points reached:
  RefBoard.get_inputs   -> deepgo::Board::getInputs           (Board.cpp:494-623)
  RefProcessor.execute  -> deepgo::Processor::execute          (Processor.cpp:36-60)
                           -> Executor::_forward -> Model::forward (Model.cpp:88-100, libtorch)
  RefProcessor.evaluate -> deepgo::Evaluator::evaluate         (Evaluator.cpp:37-84)
  RefPlayer.evaluate    -> deepgo::Player::startEvaluation/waitEvaluation/getCandidates
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / '_ref'
INPUT_SIZE = 11920
OUTPUT_SIZE = 2169
STATE_SIZE = 448
BLACK, WHITE = 1, -1
RULE_CH, RULE_JP, RULE_COM = 0, 1, 2

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags='C_CONTIGUOUS')
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags='C_CONTIGUOUS')
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags='C_CONTIGUOUS')


def available(kind: str = 'board') -> bool:
    return (REF_DIR / f'libmaru_ref_{kind}.so').exists()


_libs = {}


def _bind_board(lib) -> None:
    lib.ref_board_new.restype = C.c_void_p
    lib.ref_board_new.argtypes = [C.c_int, C.c_int]
    lib.ref_board_free.argtypes = [C.c_void_p]
    lib.ref_board_clear.argtypes = [C.c_void_p]
    lib.ref_board_play.restype = C.c_int
    lib.ref_board_play.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.ref_board_copy_from.argtypes = [C.c_void_p, C.c_void_p]
    for name in ('ref_board_get_color', 'ref_board_get_ren_space', 'ref_board_is_shicho'):
        getattr(lib, name).restype = C.c_int
        getattr(lib, name).argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.ref_board_is_enabled.restype = C.c_int
    lib.ref_board_is_enabled.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ref_board_get_enableds.argtypes = [C.c_void_p, _i32p, C.c_int, C.c_int]
    lib.ref_board_get_territories.argtypes = [C.c_void_p, _i32p, C.c_int]
    lib.ref_board_get_ko.argtypes = [C.c_void_p, C.c_int, _i32p]
    lib.ref_board_get_inputs.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, C.c_int, C.c_int]
    lib.ref_board_get_state.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_float, C.c_int, C.c_int,
                                        C.c_int]
    lib.ref_board_get_state_lean.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_float, C.c_int, C.c_int,
                                        C.c_int]


def _bind_nn(lib) -> None:
    lib.ref_processor_new.restype = C.c_void_p
    lib.ref_processor_new.argtypes = [C.c_char_p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_char_p, C.c_int]
    lib.ref_processor_free.argtypes = [C.c_void_p]
    lib.ref_processor_execute.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int]
    lib.ref_evaluator_evaluate.restype = C.c_int
    lib.ref_evaluator_evaluate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int,
                                           C.c_int, _i32p, _i32p, _f32p, _f32p]
    lib.ref_player_new.restype = C.c_void_p
    lib.ref_player_new.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                   C.c_int, C.c_int]
    lib.ref_player_free.argtypes = [C.c_void_p]
    lib.ref_player_initialize.argtypes = [C.c_void_p]
    lib.ref_player_play.restype = C.c_int
    lib.ref_player_play.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.ref_player_get_color.restype = C.c_int
    lib.ref_player_get_color.argtypes = [C.c_void_p]
    lib.ref_player_evaluate.restype = C.c_int
    lib.ref_player_evaluate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                        C.c_int, C.c_float, C.c_float, _i32p, _f32p, C.c_int]


def load(kind: str = 'board'):
    """kind: 'board' (no libtorch), 'nn' (reference libtorch runtime), 'dropin' (reference search
    and queue linked against the substitute deepgo::Model that calls libmaru_b200.so) or
    'dropin_states' (reference search + Evaluator.h with the substitute Processor / Evaluator.cpp
    that send compact maru_state records) or 'search_synth' (the same reference search objects over the
    synthetic CPU evaluator of oracle/synth_eval.c: the oracle of the native search)."""
    if kind in _libs:
        return _libs[kind]
    path = REF_DIR / f'libmaru_ref_{kind}.so'
    if not path.exists():
        raise FileNotFoundError(f'{path} missing: run `make -C oracle ref` where /root/reference exists')
    if kind == 'nn':
        import torch  # noqa: F401  libtorch must be initialised by Python first
    if kind.startswith('dropin'):
        os.environ.setdefault('MARU_B200_LIB', str(HERE.parent / 'maru_b200' / 'lib' / 'libmaru_b200.so'))
    lib = C.CDLL(str(path), mode=C.RTLD_LOCAL)
    _bind_board(lib)
    if kind != 'board':
        _bind_nn(lib)
    _libs[kind] = lib
    return lib


class RefBoard:
    """deepgo::Board (reference src/deepgo/native/cpp/Board.h) through the C shim."""

    def __init__(self, width: int = 19, height: int = 19, kind: str = 'board'):
        self.lib = load(kind)
        self.width, self.height = width, height
        self.ptr = C.c_void_p(self.lib.ref_board_new(width, height))

    def __del__(self):
        if getattr(self, 'ptr', None):
            self.lib.ref_board_free(self.ptr)
            self.ptr = None

    def play(self, x: int, y: int, color: int) -> int:
        return self.lib.ref_board_play(self.ptr, x, y, color)

    def copy(self) -> 'RefBoard':
        b = RefBoard(self.width, self.height)
        b.lib.ref_board_copy_from(b.ptr, self.ptr)
        return b

    def get_color(self, x, y) -> int:
        return self.lib.ref_board_get_color(self.ptr, x, y)

    def get_ren_space(self, x, y) -> int:
        return self.lib.ref_board_get_ren_space(self.ptr, x, y)

    def is_shicho(self, x, y) -> bool:
        return bool(self.lib.ref_board_is_shicho(self.ptr, x, y))

    def is_enabled(self, x, y, color, check_seki=False) -> bool:
        return bool(self.lib.ref_board_is_enabled(self.ptr, x, y, color, int(check_seki)))

    def get_enableds(self, color, check_seki=False) -> np.ndarray:
        out = np.zeros(self.width * self.height, np.int32)
        self.lib.ref_board_get_enableds(self.ptr, out, color, int(check_seki))
        return out

    def get_territories(self, color) -> np.ndarray:
        out = np.zeros(self.width * self.height, np.int32)
        self.lib.ref_board_get_territories(self.ptr, out, color)
        return out

    def get_ko(self, color) -> Tuple[int, int]:
        xy = np.zeros(2, np.int32)
        self.lib.ref_board_get_ko(self.ptr, color, xy)
        return int(xy[0]), int(xy[1])

    def get_inputs(self, color, komi=7.5, rule=RULE_CH, superko=False) -> np.ndarray:
        out = np.empty(INPUT_SIZE, np.float32)
        self.lib.ref_board_get_inputs(self.ptr, out, color, float(komi), rule, int(superko))
        return out

    def get_state(self, color, komi=7.5, rule=RULE_CH, superko=False, with_mask=False,
                  lean_ladder=False) -> np.ndarray:
        """Compact maru_state (448 bytes) via the Board's public getters (state_from_board.h).
        lean_ladder: ladder flags from maru_b200/dropin/lean_ladder.h instead of Board::isShicho."""
        out = np.zeros(STATE_SIZE, np.uint8)
        fn = self.lib.ref_board_get_state_lean if lean_ladder else self.lib.ref_board_get_state
        fn(self.ptr, out, color, float(komi), rule, int(superko), int(with_mask))
        return out


class RefProcessor:
    """deepgo::Processor — the reference's own queue + libtorch Model on CPU (gpus=[-1]) or CUDA."""

    def __init__(self, model: str, gpus: Sequence[int] = (-1,), batch_size: int = 2048,
                 fp16: bool = False, deterministic: bool = True, threads_par_gpu: int = 1,
                 kind: str = 'nn'):
        self.lib = load(kind)
        g = np.asarray(list(gpus), np.int32)
        err = C.create_string_buffer(512)
        self.ptr = C.c_void_p(self.lib.ref_processor_new(
            str(model).encode(), g, len(g), batch_size, int(fp16), int(deterministic),
            threads_par_gpu, err, 512))
        if not self.ptr:
            raise RuntimeError(err.value.decode())

    def close(self):
        if getattr(self, 'ptr', None):
            self.lib.ref_processor_free(self.ptr)
            self.ptr = None

    def __del__(self):
        self.close()

    def execute(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32).reshape(-1, INPUT_SIZE)
        out = np.zeros((x.shape[0], OUTPUT_SIZE), np.float32)
        self.lib.ref_processor_execute(self.ptr, x, out, x.shape[0])
        return out

    def evaluate(self, board: RefBoard, color: int, komi=7.5, rule=RULE_CH, superko=False):
        xs = np.zeros(361, np.int32)
        ys = np.zeros(361, np.int32)
        ps = np.zeros(361, np.float32)
        v = np.zeros(1, np.float32)
        n = self.lib.ref_evaluator_evaluate(self.ptr, board.ptr, color, float(komi), rule,
                                            int(superko), xs, ys, ps, v)
        return xs[:n].copy(), ys[:n].copy(), ps[:n].copy(), float(v[0])


class RefPlayer:
    """deepgo::Player — the reference MCTS on top of a RefProcessor."""

    def __init__(self, processor: RefProcessor, threads: int, width: int, height: int,
                 komi: float = 7.5, rule: int = RULE_CH, superko: bool = False,
                 eval_leaf_only: bool = False):
        self.lib = processor.lib
        self.processor = processor
        self.ptr = C.c_void_p(self.lib.ref_player_new(processor.ptr, threads, width, height,
                                                      float(komi), rule, int(superko),
                                                      int(eval_leaf_only)))

    def close(self):
        if getattr(self, 'ptr', None):
            self.lib.ref_player_free(self.ptr)
            self.ptr = None

    def __del__(self):
        self.close()

    def initialize(self):
        self.lib.ref_player_initialize(self.ptr)

    def play(self, x: int, y: int) -> int:
        return self.lib.ref_player_play(self.ptr, x, y)

    def color(self) -> int:
        return self.lib.ref_player_get_color(self.ptr)

    def evaluate(self, visits: int, playouts: int = 0, timelimit: float = 600.0,
                 equally: bool = False, use_ucb1: bool = False, width: int = 0,
                 temperature: float = 1.0, noise: float = 0.0):
        ints = np.zeros((512, 5), np.int32)
        floats = np.zeros((512, 2), np.float32)
        n = self.lib.ref_player_evaluate(self.ptr, visits, playouts, timelimit, int(equally),
                                         int(use_ucb1), width, temperature, noise, ints, floats, 512)
        return [dict(x=int(ints[i, 0]), y=int(ints[i, 1]), visits=int(ints[i, 2]),
                     playouts=int(ints[i, 3]), color=int(ints[i, 4]), policy=float(floats[i, 0]),
                     value=float(floats[i, 1]))
                for i in range(n)]


# ------------------------------------------------------------------------------------------------
# Seeded synthetic positions: random legal playouts through the reference Board.
# ------------------------------------------------------------------------------------------------
def random_playout(width: int, height: int, n_moves: int, rng: np.random.Generator,
                   pass_prob: float = 0.03) -> Tuple[RefBoard, int]:
    """Play `n_moves` random legal moves (occasional passes). Returns (board, side to move)."""
    b = RefBoard(width, height)
    color = BLACK
    for _ in range(n_moves):
        if rng.random() < pass_prob:
            b.play(-1, -1, color)
        else:
            en = b.get_enableds(color, False)
            idx = np.flatnonzero(en)
            if len(idx) == 0:
                b.play(-1, -1, color)
            else:
                # avoid filling own single-point eyes too often: keeps boards from collapsing
                for _try in range(4):
                    i = int(rng.choice(idx))
                    x, y = i % width, i // width
                    nb = [(x + dx, y + dy) for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1))
                          if 0 <= x + dx < width and 0 <= y + dy < height]
                    if not all(b.get_color(px, py) == color for px, py in nb):
                        break
                b.play(x, y, color)
        color = -color
    return b, color


def make_positions(n: int, size: int | Tuple[int, int] = 19, seed: int = 0, komi: float = 7.5,
                   with_mask: bool = False, max_fill: float = 0.6, min_moves: int = 0):
    """n seeded positions -> (rows [n,11920] fp32 from Board::getInputs, states [n,448] uint8,
    meta list). Move counts ~U[0, max_fill*w*h]; komi/rule/superko varied deterministically."""
    w, h = (size, size) if isinstance(size, int) else size
    rng = np.random.default_rng(seed)
    rows = np.zeros((n, INPUT_SIZE), np.float32)
    states = np.zeros((n, STATE_SIZE), np.uint8)
    meta = []
    for i in range(n):
        moves = int(rng.integers(min_moves, int(max_fill * w * h) + 1))
        b, color = random_playout(w, h, moves, rng, pass_prob=0.03 if min_moves == 0 else 0.0)
        rule = RULE_JP if (i % 5 == 4) else (RULE_COM if i % 7 == 6 else RULE_CH)
        superko = (i % 3 == 2)
        k = komi if i % 4 else float(rng.choice([0.5, 5.5, 6.5, 7.5, -7.5, 0.0]))
        rows[i] = b.get_inputs(color, k, rule, superko)
        states[i] = b.get_state(color, k, rule, superko, with_mask)
        meta.append(dict(w=w, h=h, moves=moves, color=color, komi=k, rule=rule, superko=superko))
    return rows, states, meta


def ko_positions(size: int | Tuple[int, int] = 9, komi: float = 7.5):
    """Forced cases the random playouts rarely hit: a ko just created (plane 31 + scalar 4, incl. the
    reference's un-centred ko plane on small boards, Board.cpp:587-593), for both colours and a few
    translations, each followed by a pass variant (ko cleared, pass absent from history)."""
    w, h = (size, size) if isinstance(size, int) else size
    rows, states, meta = [], [], []
    for (dx, dy) in [(0, 0), (w - 4, 0), (0, h - 3), (w - 4, h - 3), ((w - 4) // 2, (h - 3) // 2)]:
        for first in (BLACK, WHITE):
            b = RefBoard(w, h)
            a, o = first, -first
            for (x, y) in [(1, 0), (0, 1), (1, 2)]:
                assert b.play(dx + x, dy + y, a) >= 0
            for (x, y) in [(2, 0), (3, 1), (2, 2), (1, 1)]:
                assert b.play(dx + x, dy + y, o) >= 0
            assert b.play(dx + 2, dy + 1, a) == 1          # captures (1,1): ko for `o`
            assert b.get_ko(o) == (dx + 1, dy + 1)
            for color in (o, a):
                rows.append(b.get_inputs(color, komi, RULE_CH, False))
                states.append(b.get_state(color, komi, RULE_CH, False, True))
                meta.append(dict(w=w, h=h, ko=(dx + 1, dy + 1), color=color, case='ko'))
            b.play(-1, -1, o)                               # pass clears the ko
            rows.append(b.get_inputs(a, komi, RULE_JP, True))
            states.append(b.get_state(a, komi, RULE_JP, True, True))
            meta.append(dict(w=w, h=h, color=a, case='ko-then-pass'))
    return np.stack(rows), np.stack(states), meta
