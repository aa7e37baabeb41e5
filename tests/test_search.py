"""CPU: the native search + self-play scheduler (maru_b200/csrc/search.h, selfplay.cpp behind
maru_selfplay_*) against the UNMODIFIED reference search on the same evaluator.

Oracle = deepgo::Player / Node / NodeManager / ThreadPool compiled unchanged
(reference src/deepgo/native/cpp/Player.cpp:93-115, 180-368, Node.cpp:71-225, 460-502, 554-598) into
oracle/_ref/libmaru_ref_search_synth.so with threads = 1, driven by maru_b200.selfplay.SelfPlayScheduler
(the same schedule the native scheduler implements).  Both sides ask the deterministic synthetic
evaluator of oracle/synth_eval.c for every leaf, so with noise = 0 the two searches must agree on EVERY
move of every game and on every root statistic (visits, playouts, mean value of every candidate) — the
arithmetic of the tree policy is bit-identical, not just close.
"""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from maru_b200 import capi, selfplay

ROOT = Path(__file__).resolve().parents[1]
SYNTH = ROOT / 'oracle' / '_build' / 'liboracle_synth.so'


@pytest.fixture(scope='module')
def synth_fn(built_lib):
    if not SYNTH.exists():
        pytest.skip('oracle/_build/liboracle_synth.so missing (make -C oracle features)')
    lib = C.CDLL(str(SYNTH))
    return lib, C.cast(lib.oracle_synth_leaves, C.c_void_p)


def openings(n_games, size, n_moves, seed):
    """A few seeded legal opening moves per game (so the games differ) through the reference Board."""
    from oracle import ref
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_games):
        b, color, mv = ref.RefBoard(size, size), ref.BLACK, []
        for _ in range(n_moves):
            en = np.flatnonzero(b.get_enableds(color, True))
            i = int(rng.choice(en))
            b.play(i % size, i // size, color)
            mv.append((i % size, i // size))
            color = -color
        out.append(mv)
    return out


def reference_games(size, n_games, visits, max_moves, root, width, criterion, opens, temperature=1.0,
                    use_ucb1=False, eval_leaf_only=False):
    from oracle import ref
    proc = ref.RefProcessor('synthetic', gpus=[0], batch_size=16, kind='search_synth')
    made = []

    class Opened:
        def __init__(self, g):
            self.p = ref.RefPlayer(proc, 1, size, size, eval_leaf_only=eval_leaf_only)
            self.g = g

        def initialize(self):
            self.p.initialize()
            for (x, y) in opens[self.g]:
                self.p.play(x, y)

        def evaluate(self, **kw):
            return self.p.evaluate(**kw)

        def play(self, x, y):
            return self.p.play(x, y)

    def make(g):
        made.append(Opened(g))
        return made[-1]
    sched = selfplay.SelfPlayScheduler(make, n_games, visits, max_moves=max_moves, root=root, width=width,
                                       noise=0.0, temperature=temperature, criterion=criterion, use_ucb1=use_ucb1)
    sched.run()
    recs = sched.records
    for p in made:
        p.p.close()
    proc.close()
    return recs


def native_games(fn, size, n_games, visits, max_moves, root, width, criterion, opens, threads=2, temperature=1.0,
                 use_ucb1=False, eval_leaf_only=False):
    sp = selfplay.NativeSelfPlay(leaf_fn=fn, games=n_games, size=size, visits=visits, threads=threads,
                                 max_moves=max_moves, root=root, width=width, noise=0.0, temperature=temperature,
                                 criterion=criterion, record=True, use_ucb1=use_ucb1, eval_leaf_only=eval_leaf_only)
    for g, mv in enumerate(opens):
        for (x, y) in mv:
            assert sp.play(g, x, y) >= 0
    sp.run(max_moves)
    out = [(sp.moves(g), [sp.candidates(g, m) for m in range(len(sp.moves(g)))]) for g in range(n_games)]
    st = sp.stats()
    sp.close()
    return out, st


def key(c):
    return (c.x, c.y, c.color, c.visits, c.playouts, np.float32(c.value).tobytes())


@pytest.mark.parametrize('size,root,visits,width,criterion,temperature,ucb1,leaf_only', [
    (9, 'pucb', 40, 0, 'lcb', 1.0, False, False),
    (9, 'gumbel', 48, 8, 'lcb', 1.0, False, False),
    (13, 'pucb', 30, 0, 'lcb', 1.0, False, False),
    (9, 'pucb', 25, 0, 'lcb', 0.6, False, False),    # root temperature: pow() on the priors (Node.cpp:100-114)
    (7, 'gumbel', 33, 5, 'lcb', 1.0, False, False),
    (19, 'pucb', 30, 0, 'lcb', 1.0, False, False),
    (9, 'pucb', 40, 0, 'lcb', 1.0, True, False),      # run.py --search ucb1 (Node.cpp:211-212, 493-502)
    (9, 'pucb', 40, 0, 'lcb', 1.0, False, True),      # run.py --eval-leaf-only (Player.cpp:337-341)
])
def test_native_search_replays_the_reference_search(synth_fn, size, root, visits, width, criterion, temperature, ucb1,
                                                    leaf_only):
    from oracle import ref
    if not ref.available('search_synth') or not ref.available('board'):
        pytest.skip('oracle/_ref search_synth not present')
    _, fn = synth_fn
    n_games, max_moves = 6, 12
    opens = openings(n_games, size, 4, seed=size * 100 + visits)
    want = reference_games(size, n_games, visits, max_moves, root, width, criterion, opens, temperature, ucb1, leaf_only)
    got, st = native_games(fn, size, n_games, visits, max_moves, root, width, criterion, opens,
                           temperature=temperature, use_ucb1=ucb1, eval_leaf_only=leaf_only)
    assert st.moves == sum(len(m) for m, _ in got)
    assert st.visits == visits * st.moves                       # exactly `visits` descents per move
    compared = 0
    for g in range(n_games):
        moves, cands = got[g]
        # a reference game is comparable up to the first move where the reference overshot its visit target
        # (its stop condition is a race between two threads, see SelfPlayScheduler._play_game); rare
        upto = want[g].overshoot.index(True) if True in want[g].overshoot else len(want[g].moves)
        assert moves[:upto] == want[g].moves[:upto], f'game {g}: move sequence differs'
        for m in range(upto):
            assert sorted(map(key, cands[m])) == sorted(map(key, want[g].candidates[m])), \
                f'game {g} move {m}: root statistics differ'
        compared += upto
    assert compared >= n_games * max_moves // 2, 'the reference overshot its targets too often to compare'


def test_native_scheduler_noise_restart_and_threads(synth_fn):
    """Gumbel noise on (random, so no oracle): legal games, exact visit accounting, restarts, and the
    same games whatever the number of worker threads / lanes (a game's RNG is its own)."""
    _, fn = synth_fn
    runs = []
    for threads, lanes in ((1, 1), (3, 2), (4, 3)):
        sp = selfplay.NativeSelfPlay(leaf_fn=fn, games=10, size=9, visits=24, threads=threads, max_moves=8,
                                     root='gumbel', width=8, noise=1.0, criterion='visits', seed=7,
                                     opening_moves=3, lanes=lanes)
        sp.run(5)
        st = sp.stats()
        assert st.moves == 50 and st.visits == 24 * 50 and 0 < st.evals <= st.visits + 50
        sp.run(5)                                              # max_moves = 8: every game stops at 8
        st = sp.stats()
        assert st.moves == 80 and st.games_finished == 10
        runs.append([sp.moves(g) for g in range(10)])
        for mv in runs[-1]:
            assert len(mv) == 8 and all(-1 <= x < 9 and -1 <= y < 9 for x, y in mv)
        sp.close()
    assert runs[0] == runs[1] == runs[2]
    # restart = 1: finished games start over, the load stays constant
    sp = selfplay.NativeSelfPlay(leaf_fn=fn, games=4, size=5, visits=8, threads=2, max_moves=6, restart=True,
                                 opening_moves=2, seed=3)
    sp.run(20)
    st = sp.stats()
    assert st.moves == 80 and st.games_finished >= 4 * 3
    sp.close()


def test_selfplay_argument_errors(built_lib):
    lib = capi.load_library()
    cfg = capi.SelfPlayConfig(width=9, height=9, games=0, threads=1, visits=10, max_moves=10, temperature=1.0)
    h = C.c_void_p()
    assert lib.maru_selfplay_create(None, C.byref(cfg), C.byref(h)) != 0
    assert b'queue' in lib.maru_last_error()
    assert lib.maru_selfplay_create_fn(None, None, C.byref(cfg), C.byref(h)) != 0
