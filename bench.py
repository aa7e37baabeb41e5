#!/usr/bin/env python
"""bench.py — NN evals/sec of the leaf-evaluation hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (the reference's own CPU path, oracle/_ref)

One "step" = one pass of the hot path over one batch: `--batch` (256) synthetic 19x19 positions
through K1 encode -> trunk -> heads of the b4c128-shape random-init network (BASELINE.json
configs[1]).  Rank r runs its own evaluator on cuda:r over its own positions (positions are
independent: weak scaling, no data-path collective; NCCL only reduces the timing).
  value        : evals/s with the compact records already resident in HBM (CUDA-event timed on the
                 evaluator's stream, L2 flushed between steps, max over ranks)
  e2e          : the same metric through the blocking C-ABI call with HOST buffers
                 (maru_eval_forward_leaves: pinned compact records H2D + kernels + move priors and
                 value D2H per step — the transport of the substitute Evaluator.cpp; the full-row and
                 fp32-plane contracts are timed next to it as e2e.full_rows / e2e.planes)
  roofline     : dominant kernel (3x3 implicit-GEMM conv) algorithmic FLOPs / CUDA-event duration,
                 against the measured bf16 peak of MEASURED_PEAKS.json
  cpu_baseline : the reference runtime (deepgo::Processor on libtorch CPU) on a bounded sample
  sustained    : the same device-resident forward back to back for >= 2 s, with its own clock record
  gpu_reference: (N = 1) the UNMODIFIED reference runtime on libtorch CUDA on the same B200, fp32 / fp16
  selfplay     : the second half of the BASELINE metric, run by EVERY rank: self-play visits/s of the
                 native search + game scheduler (maru_selfplay_*: many games per host thread, leaves
                 submitted asynchronously in batches to this rank's leaf queue), games sharded by rank,
                 visits summed over ranks / slowest rank's time.  Reference arm: G reference Players
                 (deepgo::Player, unmodified) on the libtorch CPU Processor, same schedule.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = 'nn_evals_per_sec'
UNIT = 'evals/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--model', default='b4c128')
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--board', type=int, default=19)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-selfplay', action='store_true')
    ap.add_argument('--no-gpu-reference', action='store_true')
    ap.add_argument('--no-small-boards', action='store_true')
    ap.add_argument('--sustained-s', type=float, default=2.0)
    ap.add_argument('--sp-board', type=int, default=19)
    ap.add_argument('--sp-visits', type=int, default=50, help='descents per move (reference default --visits 50)')
    ap.add_argument('--sp-moves', type=int, default=8, help='timed moves per game (after 1 untimed move)')
    ap.add_argument('--sp-games', type=int, default=0,
                    help='concurrent games per GPU (0: 2560 over the queue; 480 per lane with evaluator lanes, whose batches '
                         'are not merged across threads)')
    ap.add_argument('--sp-opening', type=int, default=-1, help='random opening moves (-1: board*board/6)')
    ap.add_argument('--sp-root', default='pucb', choices=['pucb', 'gumbel'])
    ap.add_argument('--sp-max-batch', type=int, default=1024)
    ap.add_argument('--sp-backend', default='queue', choices=['lanes', 'queue'],
                    help='queue: leaves go through maru_queue (batches merged across search threads; default); lanes: the search '
                         'threads submit their own batches (maru_lane_*; measured: 510 k vs 440 k visits/s on one B200 restricted to '
                         '4 cores, but 2.74 M vs 3.08 M on the 8-GPU box whose 32 logical cores are then all taken)')
    ap.add_argument('--sp-no-pin', action='store_true', help='do not pin the search threads to this rank\'s cores')
    ap.add_argument('--sp-lanes', type=int, default=4,
                    help='leaf batches a search thread keeps in flight (4 x 2560 games: +6 % over 3 x 1920 on 8 GPUs with 3 search '
                         'threads each and on the 13x13 Gumbel leg, equal on one GPU where the evaluator is saturated either way)')
    ap.add_argument('--sp-evaluators', type=int, default=1,
                    help='evaluators (streams + buffer sets) behind the queue on this GPU: batches of two overlap on the device')
    ap.add_argument('--sp-threads', type=int, default=0,
                    help='search threads per rank (0: this rank\'s cores minus one for the queue thread)')
    return ap.parse_args()


def base_config(args, world: int) -> dict:
    """Workload description shared verbatim by both arms (identical keys and values)."""
    if (args.model, args.board, args.batch) == ('b4c128', 19, 256):
        which = ' (BASELINE.json configs[1])'
    elif args.model == 'b10c256' and args.board == 19 and args.batch in (512, 1024, 2048):
        which = ' (BASELINE.json configs[2])'
    else:
        which = ' (not a BASELINE.json config: builder-chosen variant)'
    return dict(workload=f'{args.model} random-init (seed 0), {args.board}x{args.board} leaf evaluation, '
                         f'batch {args.batch} per GPU{which}',
                model=args.model, board=args.board, batch_per_gpu=args.batch,
                parallelism=f'replica x{world}')


def selfplay_config(args) -> dict:
    opening = args.sp_opening if args.sp_opening >= 0 else args.sp_board * args.sp_board // 6
    return dict(board=args.sp_board, visits_per_move=args.sp_visits, timed_moves_per_game=args.sp_moves,
                warmup_moves_per_game=1, opening_moves=opening, root=args.sp_root, model=args.model,
                workload=f'{args.sp_board}x{args.sp_board} self-play, {args.sp_root} root, {args.sp_visits} descents per '
                         f'move on top of the inherited subtree, games sharded by rank (BASELINE.json configs[{3 if args.sp_root == "gumbel" else 4}])')


def rank_cpus(local: int) -> list:
    """This rank's share of the host cores: the process's affinity mask split evenly over the local ranks."""
    cpus = sorted(os.sched_getaffinity(0))
    nlocal = int(os.environ.get('LOCAL_WORLD_SIZE', os.environ.get('WORLD_SIZE', '1')))
    per = max(1, len(cpus) // max(1, nlocal))
    return cpus[(local % max(1, nlocal)) * per:(local % max(1, nlocal)) * per + per] or cpus


def model_files(name: str):
    """Random-init weights of the named architecture (seed 0): TorchScript .model + .b200 pack."""
    from maru_b200 import convert, netspec
    d = Path(os.environ.get('MARU_BENCH_DIR', '/tmp/maru_b200_bench'))
    d.mkdir(parents=True, exist_ok=True)
    path = d / f'{name}-seed0.model'
    if not path.exists():
        tmp = d / f'.{name}-{os.getpid()}.model'
        netspec.export(netspec.build(**netspec.CONFIGS[name], seed=0), str(tmp))
        os.replace(tmp, path)
    pack = Path(str(path) + '.b200')
    if not pack.exists():
        tmp = d / f'.{name}-{os.getpid()}.b200'
        convert.convert(str(path), str(tmp))
        os.replace(tmp, pack)
    return path, pack


def load_peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        d = json.loads(p.read_text())
        return dict(bf16_burst=d.get('bf16_tflops', 1590.0), bf16_sustained=d.get('bf16_tflops_sustained', 1400.0),
                    hbm=d.get('hbm_gbs', 6650.0), source='measured')
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source='fallback')


class ClockSampler:
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu: int):
        self.gpu, self.samples, self.stop = gpu, [], threading.Event()
        self.th = threading.Thread(target=self.run, daemon=True)

    def run_nvml(self) -> bool:
        """Fast path: poll NVML directly (~1 ms per sample) so that even a 20 ms timed region is covered by
        many samples; nvidia-smi (one process per sample, ~50 ms) is the fallback."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
        except Exception:
            return False
        bits = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}
        while not self.stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = int(get_reasons(h))
                self.samples.append([str(sm), str(mx), '0'] +
                                    ['Active' if r & b else 'Not Active' for b in (0x8, 0x40, 0x20, 0x4)])
            except Exception:
                pass
            self.stop.wait(0.001)
        return True

    def run(self):
        if self.run_nvml():
            return
        while not self.stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--id={self.gpu}', f'--query-gpu={self.QUERY}',
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for nm, v in zip(names, s[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx or None,
                    reasons=sorted(reasons), samples=len(sm))


def time_reference_cpu(model_path: Path, rows: np.ndarray, steps: int, warmup: int, budget_s: float):
    """The reference's own path on the host cores: deepgo::Processor::execute -> Executor ->
    Model::forward (libtorch CPU fp32), built unmodified into oracle/_ref.  ONE procedure for the
    reference arm and for the b200 arm's cpu_baseline leg: a step is `sample` positions of the batch, sized
    from a 16-row probe so that warmup + steps fit the budget; `warmup` untimed steps, then `steps` timed."""
    import torch
    from oracle import ref
    torch.set_num_threads(os.cpu_count() or 1)
    proc = ref.RefProcessor(str(model_path), gpus=[-1], batch_size=len(rows))
    proc.execute(rows[:16])                                    # first call: allocator / oneDNN primitive cache
    t0 = time.perf_counter()
    proc.execute(rows[:16])
    per_row = (time.perf_counter() - t0) / 16
    sample = int(max(8, min(len(rows), budget_s / (steps + warmup) / max(per_row, 1e-6))))
    for _ in range(warmup):
        proc.execute(rows[:sample])
    t0 = time.perf_counter()
    for _ in range(steps):
        proc.execute(rows[:sample])
    dt = time.perf_counter() - t0
    proc.close()
    return sample * steps / dt, sample, dt, torch.get_num_threads()


def reference_selfplay(args, model_path: Path, budget_s: float = 40.0):
    """Self-play visits/s of the UNMODIFIED reference: G deepgo::Player objects (Player.cpp, Node.cpp; each
    with its own search thread pool) over ONE deepgo::Processor on libtorch CPU, driven by the same schedule
    as the native scheduler (maru_b200.selfplay.SelfPlayScheduler).  Bounded sample: few games, 1 + 1 moves."""
    import torch
    from maru_b200 import selfplay
    from oracle import ref
    cfg = selfplay_config(args)
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    games, threads = 8, 2
    proc = ref.RefProcessor(str(model_path), gpus=[-1], batch_size=games * threads)
    made = []

    class Opened:
        def __init__(self, g):
            self.p = ref.RefPlayer(proc, threads, args.sp_board, args.sp_board)
            self.g, self.ready = g, False

        def initialize(self):
            if self.ready:
                return
            self.ready = True
            self.p.initialize()
            rng = np.random.default_rng(1000 + self.g)
            b = ref.RefBoard(args.sp_board, args.sp_board)
            color = ref.BLACK
            for _ in range(cfg['opening_moves']):
                en = np.flatnonzero(b.get_enableds(color, True))
                if len(en) == 0:
                    break
                i = int(rng.choice(en))
                b.play(i % args.sp_board, i // args.sp_board, color)
                self.p.play(i % args.sp_board, i // args.sp_board)
                color = -color

        def evaluate(self, **kw):
            return self.p.evaluate(**kw)

        def play(self, x, y):
            return self.p.play(x, y)

    def make(g):
        made.append(Opened(g))
        return made[-1]
    # probe the cost of one move with 2 games, then size the sample
    moves = 2
    sched = selfplay.SelfPlayScheduler(make, games, args.sp_visits, max_moves=moves, root=args.sp_root, width=16,
                                       noise=1.0 if args.sp_root == 'gumbel' else 0.0, criterion='lcb')
    for pl in sched.players:
        pl.initialize()
    t0 = time.perf_counter()
    st = sched.run()
    dt = time.perf_counter() - t0
    for pl in made:
        pl.p.close()
    proc.close()
    return dict(value=st.visits / dt, unit='visits/s', visits=st.visits, wall_s=dt, games=games,
                threads_per_game=threads, moves_per_game=moves, cores=os.cpu_count(),
                torch_threads=torch.get_num_threads(),
                impl='deepgo::Player x %d over deepgo::Processor on libtorch CPU fp32 (oracle/_ref, unmodified)' % games,
                config=cfg)


def rows_from_states(states: np.ndarray) -> np.ndarray:
    """CPU oracle (oracle/features.c) expansion of compact records to reference rows — used only to
    feed the reference arm / cpu_baseline leg."""
    import ctypes as C
    so = ROOT / 'oracle' / '_build' / 'liboracle_features.so'
    if not so.exists():
        subprocess.check_call(['make', '-C', str(ROOT / 'oracle'), 'features'], stdout=subprocess.DEVNULL)
    lib = C.CDLL(str(so))
    out = np.zeros((len(states), 11920), np.float32)
    lib.oracle_get_inputs_batch(states.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), len(states))
    return out


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from maru_b200 import synth
    model_path, _ = model_files(args.model)
    states = synth.random_states(args.batch, args.board, seed=1234)
    rows = rows_from_states(states)
    value, sample, dt, cores = time_reference_cpu(model_path, rows, args.steps, args.warmup, budget_s=150.0)
    sp = None
    if not args.no_selfplay:
        try:
            sp = reference_selfplay(args, model_path)
        except Exception as exc:
            sp = dict(value=None, unit='visits/s', error=str(exc))
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=dt / args.steps * 1e3, higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
                config=base_config(args, args.gpus),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind='reference',
                                  sample=f'{sample} of the {args.batch} positions per step through '
                                         'deepgo::Processor::execute (libtorch CPU fp32, oracle/_ref)'),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note='one host\'s CPU cores: this arm does not grow with --gpus (rank 0 alone runs it)',
                selfplay=sp)
    print(json.dumps(line), flush=True)


def gpu_reference_leg(model_path: Path, rows: np.ndarray, ours: np.ndarray, B: int, reps: int = 20) -> dict:
    """Same-box GPU bar: deepgo::Processor -> Executor -> Model::forward on libtorch CUDA
    (oracle/_ref, unmodified), fp32 and fp16, on the reference's own contract (fp32 rows on the host)."""
    from oracle import ref
    out = {}
    mask = rows[:, 32 * 361:33 * 361] > 0

    def cl(o):
        lp = np.where(mask, np.log(np.maximum(o[:, :361], 1e-30)), 0.0)
        return np.where(mask, lp - (lp.sum(1) / mask.sum(1))[:, None], 0.0)
    for fp16 in (False, True):
        proc = ref.RefProcessor(str(model_path), gpus=[0], batch_size=B, fp16=fp16, deterministic=False)
        for _ in range(5):                         # cudnn.benchmark tunes on the first calls of a batch size
            got = proc.execute(rows)
        t0 = time.perf_counter()
        for _ in range(reps):
            got = proc.execute(rows)
        dt = (time.perf_counter() - t0) / reps
        proc.close()
        out['fp16' if fp16 else 'fp32'] = dict(
            value=B / dt, unit=UNIT, ms_per_step=dt * 1e3,
            logit_maxabs_vs_b200=float(np.abs(cl(got) - cl(ours)).max()),
            value_maxabs_vs_b200=float(np.abs(got[:, 2166] - ours[:, 2166]).max()))
    out['api'] = ('deepgo::Processor::execute on libtorch CUDA (oracle/_ref), host fp32 rows in / out: compare with '
                  'e2e.planes (the same contract through maru_eval_forward_planes)')
    return out


def small_boards_leg(args, model_path: Path, local: int, dev, batch: int = 1024, steps: int = 10) -> dict:
    """evals/s (device-resident records, CUDA events, L2 flushed) of 9x9 and 13x13 positions next to 19x19 at the
    same batch size: small boards run on a compact frame after the stem (100 / 196 rows per position instead of the
    reference's 400; maru_eval_set_device_board tells the device entry point the size the host ones read off the
    records), so they cost less — the reference evaluates all 361 model cells for every board size."""
    import torch
    import maru_b200
    from maru_b200 import synth
    ev = maru_b200.Evaluator(model_path, gpu=local, max_batch=batch)
    stream = torch.cuda.ExternalStream(ev.stream(), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    d_out = torch.empty((batch, maru_b200.OUTPUT_SIZE), dtype=torch.float32, device=dev)
    out = dict(batch=batch, steps=steps, l2='flushed between steps (256 MiB write)')
    for board in (9, 13, 19):
        ev.set_device_board(board if board < 19 else 0, board if board < 19 else 0)
        d_states = torch.from_numpy(synth.random_states(batch, board, seed=77 + board)).to(dev)
        for _ in range(3):
            ev.forward_states_device(d_states.data_ptr(), d_out.data_ptr(), batch)
        ev.sync()
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for i in range(steps):
            with torch.cuda.stream(stream):
                flush.zero_()
                e0[i].record(stream)
            ev.forward_states_device(d_states.data_ptr(), d_out.data_ptr(), batch)
            with torch.cuda.stream(stream):
                e1[i].record(stream)
        torch.cuda.synchronize(dev)
        ms = sum(a.elapsed_time(b) for a, b in zip(e0, e1)) / steps
        out[f'{board}x{board}'] = dict(value=batch / (ms / 1e3), unit=UNIT, ms_per_step=ms,
                                       launches_per_step=ev.info().launches_per_forward)
    for b in ('9x9', '13x13'):
        out[b]['vs_19x19'] = out[b]['value'] / out['19x19']['value']
    ev.close()
    return out


def selfplay_leg(args, model_path: Path, rank: int, local: int, world: int, dist, dev) -> dict:
    """Self-play visits/s of the native scheduler on this rank's GPU; whole job = sum of visits over ranks /
    slowest rank's wall time (barrier + synchronize on both sides)."""
    import torch
    import maru_b200
    from maru_b200 import selfplay
    cfg = selfplay_config(args)
    try:
        return _selfplay_leg(args, cfg, model_path, rank, local, world, dist, dev)
    except Exception as exc:
        if dist is not None:
            raise                              # multi-rank: let torchrun report it (the other ranks are inside collectives)
        return dict(metric='selfplay_visits_per_sec', value=None, unit='visits/s', error=f'{type(exc).__name__}: {exc}',
                    config=cfg)


def _selfplay_leg(args, cfg, model_path: Path, rank: int, local: int, world: int, dist, dev) -> dict:
    import torch
    import maru_b200
    from maru_b200 import selfplay
    cpus = rank_cpus(local)
    # one core stays free for the queue's worker thread (gather, launch, scatter): it is a few % of a core, but the GPU idles
    # until it runs — measured on 4 cores: 3 search threads 429 k visits/s, 4 threads 370-410 k (pinned or not, 2-4 lanes)
    backend = args.sp_backend
    if backend == 'lanes':              # no queue thread: every core of the rank searches
        threads = args.sp_threads if args.sp_threads > 0 else max(1, len(cpus))
        n_games = args.sp_games if args.sp_games > 0 else min(480, args.sp_max_batch) * args.sp_lanes * threads
    else:
        threads = args.sp_threads if args.sp_threads > 0 else max(1, len(cpus) - 1)
        n_games = args.sp_games if args.sp_games > 0 else 2560
    games = max(threads, n_games // threads * threads)
    ev = maru_b200.Evaluator(model_path, gpu=local, max_batch=args.sp_max_batch)
    more = [maru_b200.Evaluator(model_path, gpu=local, max_batch=args.sp_max_batch)
            for _ in range(args.sp_evaluators - 1)] if backend == 'queue' else []
    q = maru_b200.LeafQueue([ev] + more, batch_size=args.sp_max_batch) if backend == 'queue' else None
    sp = selfplay.NativeSelfPlay(q, evaluator=None if q is not None else ev, games=games, size=args.sp_board, visits=args.sp_visits, threads=threads,
                                 max_moves=10 ** 6, root=args.sp_root, width=16,
                                 noise=1.0 if args.sp_root == 'gumbel' else 0.0, criterion='lcb',
                                 opening_moves=cfg['opening_moves'], restart=True, seed=1000 + rank,
                                 cpus=None if args.sp_no_pin else cpus[:threads], lanes=args.sp_lanes)

    def sync_all():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
    sp.run(1)                                                     # untimed: first move of every game
    def qstats():
        if q is not None:
            x = q.stats()
            return x.batches, x.positions, x.device_ms_total, x.max_batch_seen
        x = sp.stats()                        # lanes: one submit = one evaluator batch
        return x.submits, x.evals, x.device_s * 1e3, x.max_submit
    s0 = sp.stats()
    v0, e0, h0, st0, sub0 = s0.visits, s0.evals, s0.host_s, s0.stall_s, s0.submits
    b0, p0, d0, _ = qstats()
    sync_all()
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        sp.run(args.sp_moves)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        sync_all()
    s1 = sp.stats()
    b1, p1, d1, mb1 = qstats()
    mine = dict(visits=s1.visits - v0, evals=s1.evals - e0, wall_s=dt, host_s=s1.host_s - h0,
                stall_s=s1.stall_s - st0, submits=s1.submits - sub0, batches=b1 - b0,
                positions=p1 - p0, device_ms=d1 - d0, max_batch=mb1)
    sp.close()
    if q is not None:
        q.close()
    for e in more:
        e.close()
    ev.close()
    sums = torch.tensor([mine['visits'], mine['evals'], mine['batches'], mine['positions'], mine['host_s'],
                         mine['stall_s'], mine['device_ms']], dtype=torch.float64, device=dev)
    mx = torch.tensor([dt, float(mine['max_batch'])], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    visits, evals, batches, positions, host_s, stall_s, device_ms = [float(x) for x in sums.tolist()]
    wall = float(mx[0].item())
    gpu_busy = device_ms / 1e3 / (wall * world)
    host_busy = host_s / (wall * world * threads)
    return dict(metric='selfplay_visits_per_sec', value=visits / wall, unit='visits/s', n_gpus=world,
                evals_per_sec=evals / wall, visits=int(visits), wall_s=wall, games_per_gpu=games,
                search_threads_per_gpu=threads, host_cores_per_gpu=len(cpus),
                mean_batch=positions / max(batches, 1.0), max_batch=int(mx[1].item()),
                host_us_per_visit=host_s / max(visits, 1.0) * 1e6,
                gpu_busy_frac=gpu_busy, host_busy_frac=host_busy, stall_frac=stall_s / (wall * world * threads),
                limiter=('host cores: the search threads are saturated (per GPU: (cores - 1) / host_us_per_visit)'
                         if host_busy > max(gpu_busy, 0.85) else
                         'GPU: the evaluator is saturated' if gpu_busy >= 0.85 else
                         'neither saturated (tail of the timed region / batch latency)'),
                clocks=clocks.summary(), config=cfg,
                backend=('evaluator lanes: every search thread stages its leaves in pinned memory and submits the batch '
                         'itself (maru_lane_*)' if q is None else 'leaf-batch queue (maru_queue_submit_leaves)'),
                impl='maru_selfplay_* (native search, one descent in flight per game, asynchronous leaf batches)')


def run_b200(args):
    import torch
    import maru_b200
    from maru_b200 import synth
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)

    if rank == 0:
        model_path, pack = model_files(args.model)
    if dist is not None:
        dist.barrier()
    model_path, pack = model_files(args.model)

    B = args.batch
    ev = maru_b200.Evaluator(model_path, gpu=local, max_batch=B)
    if os.environ.get('MARU_DBG_BITS'):
        ev.debug_set(1, int(os.environ['MARU_DBG_BITS']))
    if args.board < 19:
        ev.set_device_board(args.board, args.board)    # device-resident records: the evaluator cannot read their size
    info = ev.info()
    stream = torch.cuda.ExternalStream(ev.stream(), device=dev)
    states = synth.random_states(B, args.board, seed=1234 + rank)
    d_states = torch.from_numpy(states).to(dev)
    d_out = torch.empty((B, maru_b200.OUTPUT_SIZE), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # 256 MiB > 126 MB L2
    h_states = torch.from_numpy(states).pin_memory()
    h_out = torch.empty((B, maru_b200.OUTPUT_SIZE), dtype=torch.float32).pin_memory()

    def sync_all():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- value: device-resident, CUDA events on the evaluator's stream -------------
    for _ in range(max(args.warmup, 3)):
        ev.forward_states_device(d_states.data_ptr(), d_out.data_ptr(), B)
    ev.sync()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sync_all()
    with ClockSampler(local) as clocks:
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            with torch.cuda.stream(stream):
                flush.zero_()                                         # evict L2 between timed steps
                starts[i].record(stream)
            ev.forward_states_device(d_states.data_ptr(), d_out.data_ptr(), B)
            with torch.cuda.stream(stream):
                stops[i].record(stream)
        sync_all()
        t_wall = time.perf_counter() - t_wall0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    info = ev.info()                                   # launches_per_forward of the path that actually ran
    dev_ms = float(sum(step_ms))
    clock = clocks.summary()

    # ---------------- sustained: the same forward back to back for >= 2 s (no L2 flush: one forward streams
    # ~160 MB of activations through the 126 MB L2 anyway), its own clock record ---------------------
    sustained = None
    if args.sustained_s > 0:
        n_sus = max(args.steps, int(args.sustained_s * 1e3 / max(dev_ms / args.steps, 1e-3)) + 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        with ClockSampler(local) as clocks_sus:
            with torch.cuda.stream(stream):
                e0.record(stream)
            for _ in range(n_sus):
                ev.forward_states_device(d_states.data_ptr(), d_out.data_ptr(), B)
            with torch.cuda.stream(stream):
                e1.record(stream)
            sync_all()
        sus_ms = e0.elapsed_time(e1)
        sustained = dict(forwards=n_sus, seconds=sus_ms / 1e3, ms_per_step=sus_ms / n_sus,
                         value_per_gpu=B * n_sus / (sus_ms / 1e3), clocks=clocks_sus.summary(),
                         l2='not flushed (back-to-back forwards)')

    # ---------------- e2e: blocking C-ABI call on host buffers -----------------------------------
    lib = ev.lib
    for _ in range(3):
        lib.maru_eval_forward_states(ev.handle, h_states.data_ptr(), h_out.data_ptr(), B)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rc = lib.maru_eval_forward_states(ev.handle, h_states.data_ptr(), h_out.data_ptr(), B)
        assert rc == 0
    sync_all()
    e2e_s = time.perf_counter() - t0
    # compact results out (maru_leaf: masked move priors + value scalars, what Evaluator::evaluate keeps,
    # Evaluator.cpp:55-80) — the transport of the substitute Evaluator.cpp / the leaf queue
    h_leaf = torch.empty((B, maru_b200.LEAF_SIZE), dtype=torch.float32).pin_memory()
    for _ in range(3):
        lib.maru_eval_forward_leaves(ev.handle, h_states.data_ptr(), h_leaf.data_ptr(), B)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rc = lib.maru_eval_forward_leaves(ev.handle, h_states.data_ptr(), h_leaf.data_ptr(), B)
        assert rc == 0
    sync_all()
    e2e_leaf_s = time.perf_counter() - t0
    # the same transport through the leaf-batch queue with FOUR batches in flight (asynchronous tickets: how a search
    # keeps the evaluator fed; the reference gets the same overlap from many threads blocked in Processor::execute).
    # Host buffers in and out every step, staging copies inside the timed region; reported next to the blocking number.
    e2e_async_s = float('nan')
    try:
        import ctypes as C
        q = maru_b200.LeafQueue([ev], batch_size=B)
        try:
            st_np = h_states.numpy()
            DEPTH = 4                            # tickets in flight
            lf_np = [np.zeros((B, maru_b200.LEAF_SIZE), np.float32) for _ in range(DEPTH)]
            def submit(i):
                t = C.c_void_p()
                rc = lib.maru_queue_submit_leaves(q.handle, st_np.ctypes.data, lf_np[i % DEPTH].ctypes.data, B, C.byref(t))
                assert rc == 0
                return t
            for i in range(DEPTH):
                assert lib.maru_queue_wait(q.handle, submit(i)) == 0
            sync_all()
            t0 = time.perf_counter()
            pending = []
            for i in range(args.steps):
                if len(pending) == DEPTH:
                    assert lib.maru_queue_wait(q.handle, pending.pop(0)) == 0
                pending.append(submit(i))
            for t in pending:
                assert lib.maru_queue_wait(q.handle, t) == 0
            e2e_async_s = time.perf_counter() - t0
            assert all(np.array_equal(x, h_leaf.numpy()) for x in lf_np)
        finally:
            q.close()
    except Exception as exc:                     # a side leg: never lose the line over it
        print(f'bench: queue e2e leg failed: {exc}', file=sys.stderr)
    # reference-shaped rows in (47 680 B/position), same call the substitute deepgo::Model makes
    d_rows = torch.empty((B, maru_b200.INPUT_SIZE), dtype=torch.float32, device=dev)
    ev.encode_planes_device(d_states.data_ptr(), d_rows.data_ptr(), B)
    ev.sync()
    h_rows = d_rows.cpu().pin_memory()
    for _ in range(3):
        lib.maru_eval_forward_planes(ev.handle, h_rows.data_ptr(), h_out.data_ptr(), B)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rc = lib.maru_eval_forward_planes(ev.handle, h_rows.data_ptr(), h_out.data_ptr(), B)
        assert rc == 0
    sync_all()
    e2e_planes_s = time.perf_counter() - t0

    # ---------------- per-launch profile (roofline of the dominant kernel) -----------------------
    prof = {}
    for rep in range(max(3, min(args.steps, 10))):
        for l in ev.profile_states_device(d_states.data_ptr(), d_out.data_ptr(), B):
            key = (l['kind'], l['cin'], l['cout'], l['taps'], l['name'] if l['kind'] == 5 else '')
            a = prof.setdefault(key, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0, name=l['name']))
            if rep > 0:                                            # first rep = warm-up
                a['ms'] += l['ms']; a['flops'] += l['flops']; a['bytes'] += l['bytes']; a['launches'] += 1

    # the same trunk with one launch per conv layer (debug bit 5): the 3x3 layer alone, for continuity with earlier
    # rounds and to show what the chain launch averages over (its 1x1 layers are epilogue- / latency-bound)
    per_layer = None
    if any(k[0] == 5 for k in prof):
        ev.debug_set(1, 32)
        acc = dict(ms=0.0, flops=0.0, launches=0)
        for rep in range(3):
            for l in ev.profile_states_device(d_states.data_ptr(), d_out.data_ptr(), B):
                if rep > 0 and l['kind'] == 1 and l['cin'] == l['cout'] and l['taps'] == 9:
                    acc['ms'] += l['ms']; acc['flops'] += l['flops']; acc['launches'] += 1
        ev.debug_set(1, int(os.environ.get('MARU_DBG_BITS', '0')))
        if acc['launches']:
            per_layer = acc

    ours_rows = h_out.numpy().copy()            # outputs of the plane contract (last e2e leg), for gpu_reference
    ev.close()                                  # the legs below create their own evaluators (one engine per device)

    # ---------------- gpu_reference: the unmodified reference on libtorch CUDA, same box (N = 1) -------
    gpu_ref = None
    if world == 1 and not args.no_gpu_reference:
        try:
            gpu_ref = gpu_reference_leg(model_path, h_rows.numpy(), ours_rows, B)
        except Exception as exc:
            gpu_ref = dict(error=str(exc))

    # ---------------- small boards (north_star: 9x9 / 13x13 / 19x19 throughput), N = 1 ----------------------
    small = None
    if world == 1 and not args.no_small_boards and args.board == 19:
        try:
            small = small_boards_leg(args, model_path, local, dev)
        except Exception as exc:               # a side leg must never cost the headline line
            small = dict(error=f'{type(exc).__name__}: {exc}')

    # ---------------- selfplay: every rank plays its shard of the games ---------------------------------
    sp_block = sp_13 = None
    if not args.no_selfplay:
        # (every rank takes part in the reductions inside: an exception on one rank would hang the others, so the leg
        # reduces a failure flag first)
        sp_block = selfplay_leg(args, model_path, rank, local, world, dist, dev)
        if world == 1 and sp_block is not None and 'error' not in sp_block and (args.sp_board, args.sp_root) == (19, 'pucb'):
            # BASELINE.json configs[3]: 13x13 Gumbel sequential-halving self-play, many concurrent games, 1 GPU
            a13 = argparse.Namespace(**vars(args))
            a13.sp_board, a13.sp_root, a13.sp_opening = 13, 'gumbel', -1
            sp_13 = selfplay_leg(a13, model_path, rank, local, world, dist, dev)

    # ---------------- reduce over ranks ----------------------------------------------------------
    from maru_b200 import shard
    agg = shard.reduce_stats(shard.RankStats(positions=B * args.steps, batches=args.steps,
                                             device_ms=dev_ms, wall_s=t_wall), dist, dev)
    t = torch.tensor([e2e_s, e2e_planes_s, e2e_leaf_s, e2e_async_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s, e2e_planes_s, e2e_leaf_s, e2e_async_s = [float(x) for x in t.tolist()]
    dev_ms, t_wall = agg['device_ms'], agg['wall_s']
    total = agg['positions']                      # all ranks' positions / the slowest rank's time
    value = shard.evals_per_sec(agg)

    if rank == 0:
        peaks = load_peaks()
        kinds = {0: 'encode', 1: 'conv3x3', 2: 'conv1x1', 3: 'attention', 4: 'heads', 5: 'conv flow'}
        table = []
        tot_ms = sum(a['ms'] for a in prof.values()) or 1.0
        for key, a in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
            if a['launches'] == 0:
                continue
            table.append(dict(kernel=f"{kinds[key[0]]} {key[1]}->{key[2]}" if key[0] != 5 else 'conv_flow (persistent multi-layer)',
                              example=a['name'],
                              launches_per_step=a['launches'] // max(1, (max(3, min(args.steps, 10)) - 1)),
                              us_per_launch=a['ms'] / a['launches'] * 1e3, share=a['ms'] / tot_ms,
                              tflops=a['flops'] / (a['ms'] / 1e3) / 1e12 if a['flops'] else None,
                              gbs=a['bytes'] / (a['ms'] / 1e3) / 1e9))
        # dominant kernel: the persistent conv kernel (a whole chain of 3x3 / 1x1 layers per launch) when the
        # forward uses it, else the busiest one-layer 3x3 conv
        dom_key = max((k for k in prof if k[0] in (1, 5) and prof[k]['launches']), key=lambda k: prof[k]['ms'])
        dom = prof[dom_key]
        achieved = dom['flops'] / (dom['ms'] / 1e3) / 1e12
        traffic = None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu capture
            if dom_key[0] == 5:
                capf = next(f for f in ('r2_conv_flow_ncu_full.json', 'r1_conv_flow_ncu_full.json')
                            if (ROOT / 'profiles' / f).exists())
                cap = json.loads((ROOT / 'profiles' / capf).read_text())
                traffic = cap['_traffic_bytes_per_launch'][dom['name'].split('[')[0]]
            else:
                cap = json.loads((ROOT / 'profiles' / 'r1_conv3x3_ncu_full.json').read_text())
                traffic = cap['_traffic_bytes_per_launch']['blk0.in0.c1']
        except Exception:
            pass
        roofline = dict(bound='tensor', achieved=achieved, peak=peaks['bf16_burst'], unit='TFLOP/s',
                        frac=achieved / peaks['bf16_burst'], traffic=traffic,
                        kernel=(f'conv_flow {dom["name"]} ({dom["launches"]} timed launches)' if dom_key[0] == 5 else
                                f'conv 3x3 {dom_key[1]}->{dom_key[2]} ({dom["launches"]} timed launches; conv_gemm / conv_ksplit / conv_pair by shape)'),
                        peak_source=peaks['source'] + ' (burst, kernel timed alone)',
                        flops_per_launch=dom['flops'] / dom['launches'],
                        us_per_launch=dom['ms'] / dom['launches'] * 1e3,
                        conv3x3_one_launch_per_layer=(None if per_layer is None else dict(
                            achieved=per_layer['flops'] / (per_layer['ms'] / 1e3) / 1e12,
                            frac=per_layer['flops'] / (per_layer['ms'] / 1e3) / 1e12 / peaks['bf16_burst'],
                            us_per_launch=per_layer['ms'] / per_layer['launches'] * 1e3,
                            note='same 3x3 C/2->C/2 layer timed alone on the one-launch-per-layer path (debug bit 5)')),
                        traffic_source='profiles/ ncu --set full capture of this kernel (dram__bytes_read.sum + '
                                       'dram__bytes_write.sum per launch); not re-measured in this run',
                        whole_net_tflops_per_gpu=value / world * info.flops_per_eval / 1e12,
                        whole_net_frac_of_burst_per_gpu=value / world * info.flops_per_eval / 1e12 / peaks['bf16_burst'],
                        whole_net_frac_of_sustained_per_gpu=value / world * info.flops_per_eval / 1e12 / peaks['bf16_sustained'])
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                rows = h_rows.numpy()
                v, n, dt_cpu, cores = time_reference_cpu(model_path, rows, steps=3, warmup=1, budget_s=20.0)
                cpu = dict(value=v, unit=UNIT, cores=cores, kind='reference',
                           sample=f'3 timed steps (1 warm-up) of {n} of the step\'s {B} positions through '
                                  'deepgo::Processor::execute (libtorch CPU fp32, oracle/_ref)')
            except Exception as exc:  # the oracle always exists; report instead of hiding
                cpu = dict(value=None, unit=UNIT, cores=os.cpu_count(), kind='reference',
                           sample=f'failed: {exc}')
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
            dtype='bf16', data='synthetic', impl='b200',
            config=base_config(args, world),
            details=dict(l2='flushed between steps (256 MiB write)', launches_per_step=info.launches_per_forward,
                         conv_path=('persistent multi-layer kernels (conv_flow)' if info.launches_per_forward <= 12
                                    else 'one launch per conv layer') + ', picked by the engine\'s own timing of both',
                         flops_per_eval=info.flops_per_eval, flops_per_eval_executed=info.flops_per_eval_executed),
            e2e=dict(value=total / e2e_leaf_s, unit=UNIT, h2d_bytes_per_step=B * maru_b200.STATE_SIZE,
                     d2h_bytes_per_step=B * maru_b200.LEAF_SIZE * 4,
                     api='maru_eval_forward_leaves (blocking, pinned host buffers): compact records in, '
                         'move priors + value out — the transport of the substitute Evaluator.cpp',
                     queue_four_in_flight=dict(value=(total / e2e_async_s) if e2e_async_s == e2e_async_s else None,
                                              h2d_bytes_per_step=B * maru_b200.STATE_SIZE, d2h_bytes_per_step=B * maru_b200.LEAF_SIZE * 4,
                                              api='maru_queue_submit_leaves / maru_queue_wait, four tickets in flight '
                                                  '(pageable host buffers, the queue stages them)'),
                     full_rows=dict(value=total / e2e_s, h2d_bytes_per_step=B * maru_b200.STATE_SIZE,
                                    d2h_bytes_per_step=B * maru_b200.OUTPUT_SIZE * 4,
                                    api='maru_eval_forward_states (compact records in, 2169-float rows out)'),
                     planes=dict(value=total / e2e_planes_s, h2d_bytes_per_step=B * maru_b200.INPUT_SIZE * 4,
                                 d2h_bytes_per_step=B * maru_b200.OUTPUT_SIZE * 4,
                                 api='maru_eval_forward_planes (reference Model::forward contract)')),
            gpu_launches=info.launches_per_forward * args.steps,
            clocks=clock, roofline=roofline, cpu_baseline=cpu, kernels=table,
            wall_ms_per_step=t_wall / args.steps * 1e3, sustained=sustained,
            gpu_reference=gpu_ref, small_boards=small, selfplay=sp_block, selfplay_13x13_gumbel=sp_13)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
