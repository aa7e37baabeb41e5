"""GPU: BASELINE.json configs[0] through the reference's OWN front-end on the drop-in.

oracle/_ref/frontend.tar.gz (built by `make -C oracle ref-frontend`) = the reference's run.py, deepgo/*.py
and its Cython binding modules.pyx compiled against the reference's unchanged Executor / Processor /
Evaluator / Node / Player with ONE substitution: maru_b200/dropin/Model.{h,cpp} instead of the libtorch
Model (reference src/deepgo/native/cpp/Model.h:22-41).  The tests run what a Maru user runs:

  * `python src/run.py <model> --boardsize 9 --visits 50` fed GTP `genmove b` / `genmove w`
    (reference src/run.py:101-102, src/deepgo/gtp.py:940-990) — the evaluator behind it is the B200 one;
  * the Python façade's own consumers of output planes 2-5: Player.get_territories / get_values
    (reference src/deepgo/player.py:372-463) and Player.evaluate (player.py:295-366).
"""
import os
import re
import subprocess
import sys
import tarfile
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
ARCHIVE = ROOT / 'oracle' / '_ref' / 'frontend.tar.gz'


@pytest.fixture(scope='module')
def frontend(tmp_path_factory, built_lib):
    if not ARCHIVE.exists():
        pytest.skip('oracle/_ref/frontend.tar.gz not present (make -C oracle ref-frontend)')
    d = tmp_path_factory.mktemp('frontend')
    with tarfile.open(ARCHIVE) as t:
        t.extractall(d)
    env = dict(os.environ, MARU_B200_LIB=str(built_lib), MARU_B200_MAX_BATCH='64',
               PYTHONPATH=str(d / 'src'))
    return d / 'src', env


def test_gtp_genmove_9x9_50_visits_through_run_py(frontend, model_dir):
    src, env = frontend
    model = model_dir('b4c128')
    gtp = 'boardsize 9\nclear_board\nkomi 7\ngenmove b\ngenmove w\ngenmove b\nshowboard\nquit\n'
    r = subprocess.run([sys.executable, str(src / 'run.py'), str(model), '--boardsize', '9', '--visits', '50',
                        '--threads', '4', '--initial-turn', '0', '--gpus', '0', '--batch-size', '64'],
                       input=gtp, capture_output=True, text=True, env=env, timeout=600, cwd=str(src))
    assert r.returncode == 0, r.stderr[-3000:]
    answers = [a.strip() for a in re.findall(r'^=([^\n]*)', r.stdout, flags=re.M)]
    moves = [a for a in answers if re.fullmatch(r'[A-HJ][1-9]|pass|resign', a, flags=re.I)]
    assert len(moves) == 3, (r.stdout, r.stderr[-2000:])
    assert all(m.lower() not in ('resign',) for m in moves)
    assert len({m for m in moves if m.lower() != 'pass'}) == len([m for m in moves if m.lower() != 'pass'])


def test_python_facade_consumers_of_the_output_planes(frontend, model_dir):
    src, env = frontend
    model = model_dir('b4c128')
    code = f'''
import json, numpy as np, torch
from deepgo.processor import Processor
from deepgo.player import Player
from deepgo.config import BLACK, WHITE
proc = Processor({str(model)!r}, [0], 64)
p = Player(proc, threads=4, width=9, height=9, komi=7.0)
for pos in [(2, 2), (6, 6), (2, 6), (6, 2), (4, 4)]:
    p.play(pos)
c = p.evaluate(visits=50)
terr = np.asarray(p.get_territories())
vals = np.asarray(p.get_values())
x = np.random.default_rng(0).random((3, 11920)).astype(np.float32)
x[:, 32 * 361:33 * 361] = 1.0
out = proc.execute(x)
print(json.dumps(dict(n=len(c), pos=list(c[0].pos), visits=int(sum(k.visits for k in c)),
                      terr_shape=list(terr.shape), terr_min=float(terr.min()), terr_max=float(terr.max()),
                      vals_shape=list(vals.shape), vals_min=float(vals.min()), vals_max=float(vals.max()),
                      out_shape=list(out.shape), psum=[float(v) for v in out[:, :361].sum(1)])))
'''
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, timeout=600, cwd=str(src))
    assert r.returncode == 0, r.stderr[-3000:]
    import json
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d['n'] >= 1 and 0 <= d['pos'][0] < 9 and 0 <= d['pos'][1] < 9 and d['visits'] >= 49
    assert d['terr_shape'] == [9, 9] and -1.0 <= d['terr_min'] <= d['terr_max'] <= 1.0
    assert d['vals_shape'] == [9, 9] and -1.0 <= d['vals_min'] <= d['vals_max'] <= 1.0
    assert d['out_shape'] == [3, 2169] and np.allclose(d['psum'], 1.0, atol=2e-3)
