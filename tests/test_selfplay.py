"""CPU: the self-play scheduler / Gumbel sequential-halving driver against a
scripted player that mimics the hooks of the reference search: `visits` is a target for the root's
visit count (reference Player.cpp:190, 214-216), `width` keeps the best children by value
(Node.cpp:191-198), `equally` spreads the visits over them (Node.cpp:207-210), and the played
child's subtree becomes the next root (tree reuse)."""
import numpy as np
import pytest

from maru_b200 import selfplay, shard


def test_halving_schedule_properties():
    for visits in (1, 3, 7, 50, 100, 801):
        for width in (1, 2, 5, 16, 64):
            sch = selfplay.halving_schedule(visits, width)
            assert sum(n for _, n in sch) == visits
            ws = [w for w, _ in sch]
            assert ws[0] == width and all(b == (a + 1) // 2 for a, b in zip(ws, ws[1:]))
            if visits >= len(ws) and width > 1:
                assert ws[-1] == 1 or len(sch) < int(np.ceil(np.log2(width))) + 1
            assert max(n for _, n in sch) - min(n for _, n in sch) <= 1
    assert selfplay.halving_schedule(0, 8) == []
    assert selfplay.halving_schedule(50, 16) == [(16, 10), (8, 10), (4, 10), (2, 10), (1, 10)]


class ScriptedPlayer:
    """A 1-ply 'search': candidate i has true value q[i]; every visit of a child adds one visit."""

    def __init__(self, n_moves: int, seed: int, game_len: int):
        self.rng = np.random.default_rng(seed)
        self.n_moves, self.game_len = n_moves, game_len
        self.calls = []

    def initialize(self):
        self.ply = 0
        self._new_root(0)

    def _new_root(self, inherited: int):
        self.q = self.rng.random(self.n_moves)
        self.child = np.zeros(self.n_moves, np.int64)
        self.child[0] = inherited                    # the reused subtree
        self.root_visits = inherited + (1 if inherited else 0)

    def evaluate(self, visits, equally=False, width=0, noise=0.0, temperature=1.0):
        self.calls.append((self.ply, visits, equally, width, noise))
        if self.ply >= self.game_len:
            return []
        while self.root_visits < visits:
            live = np.argsort(-self.q)[:width] if width else np.arange(self.n_moves)
            i = live[np.argmin(self.child[live])] if equally else live[0]
            self.child[i] += 1
            self.root_visits += 1
        return [dict(x=int(i % 19), y=int(i // 19), visits=int(v), value=float(self.q[i]))
                for i, v in enumerate(self.child) if v > 0]

    def play(self, x, y):
        if x >= 0:
            self._new_root(int(self.child[y * 19 + x]) - 1 if self.child[y * 19 + x] > 0 else 0)
        self.ply += 1


def test_gumbel_root_picks_the_best_and_halves():
    p = ScriptedPlayer(32, seed=1, game_len=10)
    p.initialize()
    mv, cands = selfplay.gumbel_root_search(selfplay.PlayerAdapter(p), visits=80, width=16, noise=1.0)
    total = sum(c.visits for c in cands)
    widths = [c[3] for c in p.calls]
    targets = [c[1] for c in p.calls]
    assert widths == [16, 8, 4, 2, 1] and targets == [16, 32, 48, 64, 80]
    assert all(c[2] for c in p.calls)                                  # equally=True in every phase
    assert (mv.x, mv.y) == (int(np.argmax(p.q)) % 19, int(np.argmax(p.q)) // 19)
    assert total == 80


def test_scheduler_runs_sharded_games_and_counts_new_visits():
    made = []

    def make(g):
        made.append(g)
        return ScriptedPlayer(24, seed=100 + g, game_len=6)
    world = 2
    stats = []
    for rank in range(world):
        s = selfplay.SelfPlayScheduler(make, n_games=5, visits=40, max_moves=50, root='gumbel', width=8,
                                       rank=rank, world=world)
        st = s.run()
        stats.append(st)
        for rec in s.records:
            assert rec.finished and len(rec.moves) == 8 and rec.moves[-2:] == [(-1, -1), (-1, -1)]
            # 6 searched moves of exactly 40 descents each, on top of the inherited subtree
            assert rec.visits == 6 * 40
    assert sorted(made) == [0, 1, 2, 3, 4]
    agg = shard.reduce_stats(shard.RankStats(visits=sum(s.visits for s in stats),
                                             wall_s=max(s.wall_s for s in stats)))
    assert selfplay.visits_per_sec(agg) > 0
    with pytest.raises(ValueError):
        selfplay.SelfPlayScheduler(make, 1, 10, root='nope')


def test_pucb_root_and_error_propagation():
    p = ScriptedPlayer(8, seed=3, game_len=2)
    p.initialize()
    mv, cands = selfplay.pucb_root_search(selfplay.PlayerAdapter(p), visits=20)
    assert p.calls[-1][2] is False and p.calls[-1][3] == 0 and sum(c.visits for c in cands) == 20 and mv is not None

    class Broken(ScriptedPlayer):
        def evaluate(self, **kw):
            raise RuntimeError('evaluator failed')
    s = selfplay.SelfPlayScheduler(lambda g: Broken(4, 0, 2), n_games=2, visits=5)
    with pytest.raises(RuntimeError, match='evaluator failed'):
        s.run()
