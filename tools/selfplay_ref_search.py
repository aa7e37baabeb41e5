"""Self-play throughput (visits/s, the second half of the BASELINE metric) of the UNMODIFIED reference
search on top of the B200 evaluator, driven by maru_b200.selfplay.SelfPlayScheduler.

G concurrent games in total, sharded over the ranks (game g -> rank g % world, one process per GPU:
`python -m torch.distributed.run --nproc-per-node N tools/selfplay_ref_search.py ...`). Each game is
one reference `deepgo::Player` (MCTS with `--threads` search threads) — compiled unchanged into
oracle/_ref/libmaru_ref_dropin_states.so against the substitute Processor/Evaluator of
maru_b200/dropin/states — and all games of a rank share ONE leaf-batch queue on its GPU.
--root gumbel: sequential halving at the root through the reference's equally/width/noise hooks;
--root pucb: the reference's default search. visits = root descents made for each move
(Player.cpp:300), whole job = sum over ranks / slowest rank's wall time.
With --impl reference the same driver runs on the reference's libtorch CPU Processor.

  python tools/selfplay_ref_search.py --games 32 --size 9 --visits 50 --moves 10 [--impl reference]
"""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--games', type=int, default=32, help='concurrent games over ALL ranks')
    ap.add_argument('--size', type=int, default=9)
    ap.add_argument('--visits', type=int, default=50)
    ap.add_argument('--moves', type=int, default=10)
    ap.add_argument('--threads', type=int, default=4)
    ap.add_argument('--model', default='b4c128')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--root', default='pucb', choices=['pucb', 'gumbel'])
    ap.add_argument('--width', type=int, default=16)
    ap.add_argument('--noise', type=float, default=1.0)
    ap.add_argument('--opening', type=int, default=0, help='random legal moves played before the timed search (mid-game positions)')
    args = ap.parse_args()

    import torch
    from bench import model_files
    from maru_b200 import selfplay, shard
    from oracle import ref
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('gloo')              # a handful of counters: no GPU collective needed
    if rank == 0:
        model_files(args.model)
    if dist is not None:
        dist.barrier()
    model, _ = model_files(args.model)
    kind = 'dropin_states' if args.impl == 'b200' else 'nn'
    gpus = [local] if args.impl == 'b200' else [-1]
    proc = ref.RefProcessor(str(model), gpus=gpus, batch_size=256, kind=kind)
    made = []

    import numpy as np

    class Opened:
        """A reference player that replays a seeded random legal opening when it is initialised."""

        def __init__(self, g):
            self.p = ref.RefPlayer(proc, args.threads, args.size, args.size)
            self.g = g
            self.ready = False

        def initialize(self):
            if self.ready:               # the opening is replayed once, outside the timed region
                return
            self.ready = True
            self.p.initialize()
            if args.opening > 0:
                rng = np.random.default_rng(1000 + self.g)
                b = ref.RefBoard(args.size, args.size)
                color = ref.BLACK
                for _ in range(args.opening):
                    en = np.flatnonzero(b.get_enableds(color, True))
                    if len(en) == 0:
                        break
                    i = int(rng.choice(en))
                    x, y = i % args.size, i // args.size
                    b.play(x, y, color)
                    self.p.play(x, y)
                    color = -color

        def evaluate(self, **kw):
            return self.p.evaluate(**kw)

        def play(self, x, y):
            return self.p.play(x, y)

        def close(self):
            self.p.close()

    def make_player(g):
        p = Opened(g)
        made.append(p)
        return p
    sched = selfplay.SelfPlayScheduler(make_player, args.games, args.visits, max_moves=args.moves,
                                       root=args.root, width=args.width, noise=args.noise,
                                       rank=rank, world=world)
    for p in sched.players:
        p.initialize()
    if dist is not None:
        dist.barrier()
    st = sched.run()
    agg = shard.reduce_stats(st, dist)
    if rank == 0:
        print(json.dumps(dict(metric='selfplay_visits_per_sec', value=selfplay.visits_per_sec(agg),
                              unit='visits/s', n_gpus=world if args.impl == 'b200' else 0, impl=args.impl,
                              games=args.games, board=args.size, visits_per_move=args.visits,
                              moves_per_game=args.moves, threads_per_game=args.threads, root=args.root,
                              opening_moves=args.opening, lean_ladder=os.environ.get('MARU_LEAN_LADDER', '1') != '0',
                              total_visits=agg['visits'], moves=agg['batches'], wall_s=agg['wall_s'],
                              host_cores=os.cpu_count(), model=args.model)), flush=True)
    for p in made:
        p.close()
    proc.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
