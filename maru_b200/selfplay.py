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

    @property
    def is_pass(self) -> bool:
        return self.x < 0 or self.y < 0


class PlayerAdapter:
    """Uniform view of a reference-style player (see module docstring)."""

    def __init__(self, player):
        self.player = player

    def initialize(self) -> None:
        self.player.initialize()

    def evaluate(self, **kw) -> List[Move]:
        out = []
        for c in self.player.evaluate(**kw):
            if isinstance(c, dict):
                out.append(Move(int(c['x']), int(c['y']), int(c['visits']), float(c['value'])))
            else:                                    # reference deepgo.player.Candidate
                out.append(Move(int(c.pos[0]), int(c.pos[1]), int(c.visits), float(c.value)))
        return out

    def play(self, mv: Move) -> None:
        try:
            self.player.play((mv.x, mv.y))           # reference Python Player.play(pos)
        except TypeError:
            self.player.play(mv.x, mv.y)             # native-shim style play(x, y)


def _lcb(mv: Move, color_sign: float = 1.0) -> float:
    """Candidate ordering of the reference (player.py:92-93, 349-352): win chance LCB."""
    win = mv.value * color_sign * 0.5 + 0.5
    return win - 1.96 * 0.25 / math.sqrt(mv.visits + 1)


def gumbel_root_search(player: PlayerAdapter, visits: int, width: int = 16, noise: float = 1.0,
                       temperature: float = 1.0, criterion: str = 'visits') -> Tuple[Optional[Move], int]:
    """One move decision by sequential halving at the root.
    Returns (move or None, visits now held by the root's children).

    Phase k calls evaluate(visits=n_1+...+n_k, equally=True, width=m_k, noise=noise): the reference's
    `visits` is a target for the ROOT's visit count (Player.cpp:190, 214-216) and the tree persists
    between calls (startEvaluation only re-arms the search, Player.cpp:180-203), so phase k continues
    from phase k-1 and Node::evaluate's LCB cut (Node.cpp:191-198) drops the worse half of the
    survivors."""
    cands: List[Move] = []
    target = 0
    for w, n in halving_schedule(visits, width):
        target += n
        cands = player.evaluate(visits=target, equally=True, width=w, noise=noise, temperature=temperature)
        if not cands:
            break
    if not cands:
        return None, 0
    best = max(cands, key=(lambda c: c.visits) if criterion == 'visits' else _lcb)
    return best, sum(c.visits for c in cands)


def pucb_root_search(player: PlayerAdapter, visits: int, temperature: float = 1.0,
                     criterion: str = 'visits') -> Tuple[Optional[Move], int]:
    """The reference's default search (PUCB everywhere, run.py / gtp.py genmove)."""
    cands = player.evaluate(visits=visits, equally=False, width=0, noise=0.0, temperature=temperature)
    if not cands:
        return None, 0
    best = max(cands, key=(lambda c: c.visits) if criterion == 'visits' else _lcb)
    return best, sum(c.visits for c in cands)


@dataclass
class GameRecord:
    moves: List[Tuple[int, int]] = field(default_factory=list)
    visits: int = 0
    finished: bool = False


class SelfPlayScheduler:
    """Runs this rank's shard of `n_games` concurrent self-play games.

    make_player(game_index) -> a reference-style player bound to this rank's leaf-batch queue."""

    def __init__(self, make_player: Callable[[int], object], n_games: int, visits: int,
                 max_moves: int = 400, root: str = 'gumbel', width: int = 16, noise: float = 1.0,
                 temperature: float = 1.0, criterion: str = 'visits', rank: int = 0, world: int = 1):
        if root not in ('gumbel', 'pucb'):
            raise ValueError(f'unknown root search {root!r}')
        self.games = shard.shard_items(n_games, rank, world)
        self.players = [PlayerAdapter(make_player(g)) for g in self.games]
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
            carried = 0          # visits the new root inherited from the previous search (tree reuse)
            for _ in range(self.max_moves):
                if self.root == 'gumbel':
                    mv, spent = gumbel_root_search(p, self.visits, self.width, self.noise,
                                                   self.temperature, self.criterion)
                else:
                    mv, spent = pucb_root_search(p, self.visits, self.temperature, self.criterion)
                rec.visits += max(0, spent - carried)    # only the descents made for THIS move
                carried = max(0, mv.visits - 1) if mv is not None else 0
                if mv is None:
                    mv = Move(PASS[0], PASS[1], 0, 0.0)
                rec.moves.append((mv.x, mv.y))
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


def visits_per_sec(agg) -> float:
    return agg['visits'] / agg['wall_s'] if agg['wall_s'] > 0 else 0.0
