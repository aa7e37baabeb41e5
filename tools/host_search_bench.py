"""Host-only throughput of the native search + scheduler over the synthetic evaluator (oracle/_build/liboracle_synth.so):
usage: host_search_bench.py <board> <threads> <games> <opening moves> [pucb|gumbel]"""
import sys, time; sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
import ctypes as C
from maru_b200 import selfplay
lib = C.CDLL('' + str(__import__('pathlib').Path(__file__).resolve().parents[1] / 'oracle' / '_build' / 'liboracle_synth.so') + ''); fn = C.cast(lib.oracle_synth_leaves, C.c_void_p)
size=int(sys.argv[1]); threads=int(sys.argv[2]); games=int(sys.argv[3]); opening=int(sys.argv[4]); root=sys.argv[5] if len(sys.argv)>5 else 'pucb'
sp = selfplay.NativeSelfPlay(leaf_fn=fn, games=games, size=size, visits=50, threads=threads, max_moves=400, root=root, width=16, noise=1.0 if root=='gumbel' else 0.0, criterion='lcb', opening_moves=opening, restart=True, seed=1)
sp.run(1)
t=time.perf_counter(); s0=sp.stats(); v0,e0,h0=s0.visits,s0.evals,s0.host_s
sp.run(6); dt=time.perf_counter()-t; st=sp.stats()
print(f'{size}x{size} threads={threads} games={games} opening={opening} {root}: {(st.visits-v0)/dt:.0f} visits/s, {(st.evals-e0)/dt:.0f} evals/s, host {(st.host_s-h0)/(st.visits-v0)*1e6:.1f} us/visit, stall {st.stall_s:.2f}s')
print('expansions/evals', st.expansions/st.evals)
