"""Self-play throughput of the UNMODIFIED reference search on top of the B200 evaluator.

G concurrent games, each one reference `deepgo::Player` (MCTS, `threads` search threads) — compiled
unchanged into oracle/_ref/libmaru_ref_dropin_states.so against the substitute Processor/Evaluator of
maru_b200/dropin/states — all sharing ONE leaf-batch queue per GPU. Every move: Player.evaluate(visits=V)
then play the most visited candidate. Reports visits/s (root descents, Player.cpp:300) and the queue's
batching statistics. With --impl reference the same driver runs on the reference's libtorch CPU
Processor (oracle/_ref/libmaru_ref_nn.so) for comparison.

  python tools/selfplay_ref_search.py --games 32 --size 9 --visits 50 --moves 10 [--impl reference]
"""
import argparse
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--games', type=int, default=32)
    ap.add_argument('--size', type=int, default=9)
    ap.add_argument('--visits', type=int, default=50)
    ap.add_argument('--moves', type=int, default=10)
    ap.add_argument('--threads', type=int, default=4)
    ap.add_argument('--model', default='b4c128')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--gpu', type=int, default=0)
    args = ap.parse_args()

    import torch  # noqa: F401
    from bench import model_files
    from oracle import ref
    model, _ = model_files(args.model)
    kind = 'dropin_states' if args.impl == 'b200' else 'nn'
    gpus = [args.gpu] if args.impl == 'b200' else [-1]
    proc = ref.RefProcessor(str(model), gpus=gpus, batch_size=256, kind=kind)
    players = [ref.RefPlayer(proc, args.threads, args.size, args.size) for _ in range(args.games)]
    for p in players:
        p.initialize()
    visits = [0] * args.games

    def game(i):
        p = players[i]
        for _ in range(args.moves):
            cands = p.evaluate(visits=args.visits)
            if not cands:
                break
            visits[i] += sum(c['visits'] for c in cands)
            best = max(cands, key=lambda c: c['visits'])
            p.play(best['x'], best['y'])

    t0 = time.perf_counter()
    ths = [threading.Thread(target=game, args=(i,)) for i in range(args.games)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    dt = time.perf_counter() - t0
    total = sum(visits)
    print(f'impl={args.impl} games={args.games} size={args.size} visits/move={args.visits} moves={args.moves} '
          f'threads/game={args.threads}: {total} visits in {dt:.2f} s = {total / dt:.0f} visits/s')
    for p in players:
        p.close()
    proc.close()


if __name__ == '__main__':
    main()
