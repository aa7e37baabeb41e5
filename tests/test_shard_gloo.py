"""CPU, world_size 2 over gloo: the N>1 path of the benchmark / self-play driver — static game
sharding and the reduction of per-rank counters (sum) and times (max over ranks)."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def _free_port() -> int:
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank: int, world: int, port: int, out):
    import torch.distributed as dist
    from maru_b200 import shard
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    games = shard.shard_items(11, rank, world)
    st = shard.RankStats(positions=256 * len(games), visits=50 * len(games), batches=len(games),
                         device_ms=10.0 * (rank + 1), wall_s=0.5 + rank)
    agg = shard.reduce_stats(st, dist)
    out[rank] = (games, agg)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_and_reduce():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    g0, a0 = out[0]
    g1, a1 = out[1]
    assert sorted(g0 + g1) == list(range(11)) and not set(g0) & set(g1)
    assert g0 == [0, 2, 4, 6, 8, 10] and g1 == [1, 3, 5, 7, 9]
    assert a0 == a1                                   # every rank sees the same aggregate
    assert a0['positions'] == 256 * 11 and a0['visits'] == 50 * 11 and a0['batches'] == 11
    assert a0['device_ms'] == 20.0 and a0['wall_s'] == 1.5 and a0['world'] == 2
    from maru_b200 import shard
    assert shard.evals_per_sec(a0) == pytest.approx(256 * 11 / 0.020)


def _selfplay_worker(rank: int, world: int, port: int, out):
    """One rank of the self-play job: its shard of the games on the native scheduler (synthetic evaluator: no GPU here),
    then the reduction bench.py does — visits summed, wall time = slowest rank."""
    import ctypes as C
    import time
    import torch.distributed as dist
    from maru_b200 import selfplay, shard
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lib = C.CDLL(str(ROOT / 'oracle' / '_build' / 'liboracle_synth.so'))
    fn = C.cast(lib.oracle_synth_leaves, C.c_void_p)
    games = shard.shard_items(10, rank, world)
    sp = selfplay.NativeSelfPlay(leaf_fn=fn, games=len(games), size=9, visits=16, threads=2, max_moves=5, root='gumbel',
                                 width=4, noise=1.0, criterion='visits', seed=100 + rank, opening_moves=2)
    dist.barrier()
    t0 = time.perf_counter()
    sp.run(5)
    st = sp.stats()
    agg = shard.reduce_stats(shard.RankStats(visits=st.visits, positions=st.evals, batches=st.moves,
                                             wall_s=time.perf_counter() - t0), dist)
    out[rank] = (len(games), st.visits, st.moves, agg)
    sp.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_play_their_shards_of_the_games():
    if not (ROOT / 'oracle' / '_build' / 'liboracle_synth.so').exists() or \
            not (ROOT / 'maru_b200' / 'lib' / 'libmaru_b200.so').exists():
        pytest.skip('native libraries not built')
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_selfplay_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    (n0, v0, m0, a0), (n1, v1, m1, a1) = out[0], out[1]
    assert n0 + n1 == 10 and m0 == 5 * n0 and m1 == 5 * n1 and v0 == 16 * m0 and v1 == 16 * m1
    assert a0['visits'] == a1['visits'] == v0 + v1 == 16 * 50
    assert a0['wall_s'] == a1['wall_s'] > 0
    from maru_b200 import selfplay
    assert selfplay.visits_per_sec(a0) > 0


def test_single_process_passthrough():
    from maru_b200 import shard
    agg = shard.reduce_stats(shard.RankStats(positions=5, device_ms=1.0))
    assert agg['positions'] == 5 and agg['world'] == 1
    with pytest.raises(ValueError):
        shard.shard_items(4, 2, 2)
