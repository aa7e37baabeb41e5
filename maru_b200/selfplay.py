"""Self-play game scheduler with a Gumbel sequential-halving root.

The reference ships the HOOKS of a Gumbel root search but not the schedule that drives them:
`Node::evaluate(equally, width, useUcb1, temperature, noise)` (reference
src/deepgo/native/cpp/Node.cpp:71-225) adds Gumbel(0, noise) to the root log-priors when it picks
the next candidate to register (:99-140), keeps only the `width` best children by value LCB
(:191-198) and, with `equally`, spreads the visits evenly over them (:207-210); the Python
`Player.evaluate(visits, ..., equally, width, noise)` (reference src/deepgo/player.py:295-327)
passes them through.  This module is the missing driver:

  * `halving_schedule(visits, width)` — the sequential-halving phases (width m, m/2, ..., 1 with
    the visit budget split evenly over the phases);
  * `gumbel_root_search(player, ...)` — one move decision: one `evaluate` call per phase, each
    continuing the same search tree with half the root width of the previous one;
  * `SelfPlayScheduler` — G concurrent games (one host thread per game; the search threads live
    inside the native Player, the GIL is released during `evaluate`), all feeding ONE leaf-batch
    queue per GPU, so the evaluator sees batches of up to G x threads leaves instead of <= threads.

Multi-GPU: games are sharded statically (game g -> rank g % world, `maru_b200.shard`), one process
per GPU, no data-path collective; only the counters are reduced at report time.

`NativeSelfPlay` is the product scheduler: the same move decision and the same tree policy in C++
(maru_b200/csrc/search.h, selfplay.cpp behind `maru_selfplay_*`), hundreds of games per host thread,
leaves submitted asynchronously in batches.  `SelfPlayScheduler` + `PlayerAdapter` below drive ANY
reference-style player (the reference's own `deepgo::Player` in the tests and in bench's reference
arm) through the identical schedule, which is what the native scheduler is checked against.

Players are duck-typed after the reference's `deepgo.player.Player`:
    initialize(); evaluate(visits=, equally=, width=, noise=, temperature=) -> candidates; play(move)
A candidate is either a reference `Candidate` (`.pos`, `.visits`, `.value`, `.color`) or a mapping
with keys x, y, visits, value.  `PlayerAdapter` hides the difference.
"""
from __future__ import annotations

import math
import threading
import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

from . import shard

PASS = (-1, -1)


def halving_schedule(visits: int, width: int) -> List[Tuple[int, int]]:
    """[(root width, visits of the phase), ...] of sequential halving over `width` candidates.

    ceil(log2(width)) + 1 phases with widths m, ceil(m/2), ..., 1; the budget is split evenly over
    the phases (earlier phases receive the remainder), every phase gets at least one visit per
    surviving candidate as long as the budget allows, and the phase visits always sum to `visits`."""
    if visits <= 0:
        return []
    width = max(1, int(width))
    widths = [width]
    while widths[-1] > 1:
        widths.append((widths[-1] + 1) // 2)
    n = len(widths)
    if visits < n:                                   # fewer visits than phases: widest phases first
        return [(w, 1) for w in widths[:visits]]
    base, rem = divmod(visits, n)
    return [(w, base + (1 if i < rem else 0)) for i, w in enumerate(widths)]


@dataclass
class Move:
    x: int
    y: int
    visits: int
    value: float
    color: int = 1
    playouts: int = 0
    policy: float = 0.0

    @property
    def is_pass(self) -> bool:
        return self.x < 0 or self.y < 0


class PlayerAdapter:
    """Uniform view of a reference-style player (see module docstring)."""

    def __init__(self, player, **extra):
        self.player = player
        self.extra = extra           # fixed keyword arguments of every evaluate() call (use_ucb1=...)

    def initialize(self) -> None:
        self.player.initialize()

    def evaluate(self, **kw) -> List[Move]:
        kw = dict(kw, **self.extra)
        out = []
        for c in self.player.evaluate(**kw):
            if isinstance(c, dict):
                out.append(Move(int(c['x']), int(c['y']), int(c['visits']), float(c['value']),
                                int(c.get('color', 1)), int(c.get('playouts', 0)), float(c.get('policy', 0.0))))
            else:                                    # reference deepgo.player.Candidate
                out.append(Move(int(c.pos[0]), int(c.pos[1]), int(c.visits), float(c.value),
                                int(c.color), int(c.playouts), float(c.policy)))
        return out

    def play(self, mv: Move) -> None:
        try:
            self.player.play((mv.x, mv.y))           # reference Python Player.play(pos)
        except TypeError:
            self.player.play(mv.x, mv.y)             # native-shim style play(x, y)


def _lcb(mv: Move) -> float:
    """Candidate ordering of the reference (player.py:92-93, 349-352): LCB of the MOVER's win chance
    (value is black's point of view, color the colour of the candidate stone)."""
    win = mv.value * mv.color * 0.5 + 0.5
    return win - 1.96 * 0.25 / math.sqrt(mv.visits + 1)


def _best(cands: List[Move], criterion: str) -> Move:
    return max(cands, key=(lambda c: c.visits) if criterion == 'visits' else _lcb)   # first of equals


def gumbel_root_search(player: PlayerAdapter, visits: int, width: int = 16, noise: float = 1.0,
                       temperature: float = 1.0, criterion: str = 'visits', base: int = 0
                       ) -> Tuple[Optional[Move], List[Move]]:
    """One move decision by sequential halving at the root.  Returns (move, candidate list).

    The reference's `visits` is a target for the ROOT's visit count (Player.cpp:189, 214-219) and the
    tree persists between calls (startEvaluation only re-arms the search, Player.cpp:180-203).  `base`
    is the visit count the root inherited with its subtree, so phase k asks for
    base + n_1 + ... + n_k: exactly `visits` descents per move, phase k continuing phase k-1's tree
    with equally=True and half the root width — Node::evaluate's LCB cut (Node.cpp:191-198) drops the
    worse half of the survivors."""
    cands: List[Move] = []
    target = base
    for w, n in halving_schedule(visits, width):
        target += n
        cands = player.evaluate(visits=target, equally=True, width=w, noise=noise, temperature=temperature)
        if not cands:
            break
    if not cands:
        return None, []
    return _best(cands, criterion), cands


def pucb_root_search(player: PlayerAdapter, visits: int, temperature: float = 1.0,
                     criterion: str = 'visits', base: int = 0) -> Tuple[Optional[Move], List[Move]]:
    """The reference's default search (PUCB everywhere, run.py / gtp.py genmove): `visits` descents on
    top of the `base` visits the root inherited."""
    cands = player.evaluate(visits=base + visits, equally=False, width=0, noise=0.0, temperature=temperature)
    if not cands:
        return None, []
    return _best(cands, criterion), cands


@dataclass
class GameRecord:
    moves: List[Tuple[int, int]] = field(default_factory=list)
    candidates: List[List[Move]] = field(default_factory=list)
    overshoot: List[bool] = field(default_factory=list)   # the player made more descents than asked (see _play_game)
    visits: int = 0
    finished: bool = False


class SelfPlayScheduler:
    """Runs this rank's shard of `n_games` concurrent self-play games over reference-style players
    (one host thread per game; the oracle / reference arm of the native scheduler below).

    make_player(game_index) -> a reference-style player bound to this rank's leaf-batch queue."""

    def __init__(self, make_player: Callable[[int], object], n_games: int, visits: int,
                 max_moves: int = 400, root: str = 'gumbel', width: int = 16, noise: float = 1.0,
                 temperature: float = 1.0, criterion: str = 'visits', rank: int = 0, world: int = 1,
                 use_ucb1: bool = False):
        if root not in ('gumbel', 'pucb'):
            raise ValueError(f'unknown root search {root!r}')
        self.games = shard.shard_items(n_games, rank, world)
        self.players = [PlayerAdapter(make_player(g), **({'use_ucb1': True} if use_ucb1 else {})) for g in self.games]
        self.visits, self.max_moves = visits, max_moves
        self.root, self.width, self.noise = root, width, noise
        self.temperature, self.criterion = temperature, criterion
        self.records = [GameRecord() for _ in self.games]
        self.errors: List[BaseException] = []

    def _play_game(self, i: int) -> None:
        p, rec = self.players[i], self.records[i]
        try:
            p.initialize()
            passes = 0
            base = 0             # visits the root inherited with its subtree (tree reuse, Player.cpp:104)
            for _ in range(self.max_moves):
                if self.root == 'gumbel':
                    mv, cands = gumbel_root_search(p, self.visits, self.width, self.noise,
                                                   self.temperature, self.criterion, base)
                else:
                    mv, cands = pucb_root_search(p, self.visits, self.temperature, self.criterion, base)
                rec.visits += self.visits if cands else 0    # descents made for THIS move
                if mv is None:
                    mv = Move(PASS[0], PASS[1], 0, 0.0)
                # the chosen child becomes the root and brings its visit count; the policy / pass
                # fallback of a childless root (ONE candidate reported with visits = playouts = 1 and
                # policy 1.0, Player.cpp:246-254) starts from a fresh node instead
                fallback = len(cands) <= 1 and (not cands or (cands[0].visits == 1 and cands[0].playouts == 1
                                                              and cands[0].policy == 1.0))
                # The reference only stops when its waiter thread has seen the target (Player.cpp:212-223) while
                # its control thread keeps launching descents (Player.cpp:282-313): with a very fast evaluator
                # one extra descent can slip in.  Recorded so that comparisons can tell (root visits = 1 + sum
                # of the children's visits).
                rec.overshoot.append(bool(cands) and not fallback and
                                     sum(c.visits for c in cands) + 1 != base + self.visits)
                base = 0 if fallback else mv.visits
                rec.moves.append((mv.x, mv.y))
                rec.candidates.append(cands)
                p.play(mv)
                passes = passes + 1 if mv.is_pass else 0
                if passes >= 2:
                    rec.finished = True
                    break
        except BaseException as exc:  # surface worker failures to run()
            self.errors.append(exc)

    def run(self) -> shard.RankStats:
        t0 = time.perf_counter()
        ths = [threading.Thread(target=self._play_game, args=(i,), name=f'game{g}')
               for i, g in enumerate(self.games)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        if self.errors:
            raise self.errors[0]
        return shard.RankStats(visits=sum(r.visits for r in self.records),
                               batches=sum(len(r.moves) for r in self.records),
                               wall_s=time.perf_counter() - t0)


class NativeSelfPlay:
    """The product scheduler: `maru_selfplay_*` (maru_b200/csrc/selfplay.cpp, search.h).

    queue: a `LeafQueue` on this rank's GPU; or evaluator=: an `Evaluator` used directly through lanes (the search threads
    stage their leaves in pinned memory and submit them themselves: no queue thread, every core of the rank can search);
    or leaf_fn=: a C function pointer of type maru_leaf_fn (another evaluator: the CPU tests pass the oracle's synthetic
    one).  Games are this rank's shard."""

    ROOTS = {'pucb': 0, 'gumbel': 1}
    CRITERIA = {'lcb': 0, 'visits': 1}

    def __init__(self, queue=None, *, evaluator=None, leaf_fn=None, games: int, size: int | Tuple[int, int] = 19, visits: int = 50,
                 threads: int = 1, max_moves: int = 400, root: str = 'pucb', width: int = 16, noise: float = 0.0,
                 temperature: float = 1.0, criterion: str = 'lcb', komi: float = 7.5, rule: int = 0,
                 superko: bool = False, eval_leaf_only: bool = False, opening_moves: int = 0,
                 restart: bool = False, seed: int = 0, cpus: Optional[Sequence[int]] = None, lanes: int = 2,
                 record: bool = False, use_ucb1: bool = False):
        import ctypes as C
        from . import capi
        self.lib = capi.load_library()
        w, h = (size, size) if isinstance(size, int) else size
        cfg = capi.SelfPlayConfig(
            width=w, height=h, komi=komi, rule=rule, superko=int(superko), games=games, threads=threads,
            visits=visits, max_moves=max_moves, root=self.ROOTS[root], root_width=width, noise=noise,
            temperature=temperature, criterion=self.CRITERIA[criterion], eval_leaf_only=int(eval_leaf_only),
            opening_moves=opening_moves, restart=int(restart), seed=seed,
            cpu_first=(min(cpus) if cpus else 0), cpu_count=(len(cpus) if cpus else 0), lanes=lanes,
            record=int(record), use_ucb1=int(use_ucb1))
        self.games = games
        self._keep = (queue, evaluator, leaf_fn)
        h_ = C.c_void_p()
        if evaluator is not None:       # evaluator lanes: the search threads submit their own batches (no queue thread)
            rc = self.lib.maru_selfplay_create_eval(evaluator.handle, C.byref(cfg), C.byref(h_))
        elif queue is not None:
            rc = self.lib.maru_selfplay_create(queue.handle, C.byref(cfg), C.byref(h_))
        else:
            rc = self.lib.maru_selfplay_create_fn(C.cast(leaf_fn, C.c_void_p), None, C.byref(cfg), C.byref(h_))
        capi._check(self.lib, rc, 'maru_selfplay_create')
        self.handle = h_

    def close(self) -> None:
        if getattr(self, 'handle', None):
            self.lib.maru_selfplay_destroy(self.handle)
            self.handle = None

    def __del__(self):
        self.close()

    def play(self, game: int, x: int, y: int) -> int:
        return self.lib.maru_selfplay_play(self.handle, game, x, y)

    def run(self, moves: int) -> None:
        """Every unfinished game searches and plays `moves` more moves (blocking)."""
        from . import capi
        capi._check(self.lib, self.lib.maru_selfplay_run(self.handle, moves), 'maru_selfplay_run')

    def stats(self):
        import ctypes as C
        from . import capi
        st = capi.SelfPlayStats()
        self.lib.maru_selfplay_get_stats(self.handle, C.byref(st))
        return st

    def moves(self, game: int) -> List[Tuple[int, int]]:
        import numpy as np
        n = self.lib.maru_selfplay_get_moves(self.handle, game, None, 0)
        xy = np.zeros((max(n, 1), 2), np.int16)
        self.lib.maru_selfplay_get_moves(self.handle, game, xy.ctypes.data, n)
        return [(int(xy[i, 0]), int(xy[i, 1])) for i in range(n)]

    def candidates(self, game: int, move: int) -> List[Move]:
        from . import capi
        arr = (capi.CandidateRec * 400)()
        n = self.lib.maru_selfplay_get_candidates(self.handle, game, move, arr, 400)
        return [Move(arr[i].x, arr[i].y, arr[i].visits, arr[i].value, arr[i].color, arr[i].playouts, arr[i].prior)
                for i in range(max(n, 0))]


def visits_per_sec(agg) -> float:
    return agg['visits'] / agg['wall_s'] if agg['wall_s'] > 0 else 0.0
