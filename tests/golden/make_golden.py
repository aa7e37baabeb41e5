"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref), in the container that
has /root/reference. Run once; the fixtures travel to the GPU box, the reference sources do not.

  positions_<tag>.npz : seeded positions as compact records + the rows deepgo::Board::getInputs
                        (Board.cpp:494-623) produced for them (planes bit-packed, scalars fp32)
  outputs_b2c64.npz   : deepgo::Processor::execute (CPU fp32 libtorch, Model.cpp:88-100) of the
                        netspec b2c64 seed-0 TorchScript export on those rows

usage: python tests/golden/make_golden.py
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ref  # noqa: E402

HERE = Path(__file__).resolve().parent
SETS = [('9', 9, 24, 11), ('13', 13, 24, 12), ('19', 19, 24, 13), ('9x13', (9, 13), 16, 14),
        ('5x7', (5, 7), 8, 15)]


def pack_rows(rows: np.ndarray):
    planes = (rows[:, :33 * 361] != 0)
    assert np.array_equal(planes.astype(np.float32), rows[:, :33 * 361]), 'planes must be 0/1'
    return np.packbits(planes, axis=1), rows[:, 33 * 361:].copy()


def main():
    all_rows = []
    for tag, size, n, seed in SETS:
        rows, states, meta = ref.make_positions(n, size, seed=seed, with_mask=True)
        krows, kstates, kmeta = ref.ko_positions(size)
        rows = np.concatenate([rows, krows])
        states = np.concatenate([states, kstates])
        bits, scal = pack_rows(rows)
        np.savez_compressed(HERE / f'positions_{tag}.npz', states=states, plane_bits=bits,
                            scalars=scal, n_random=n)
        all_rows.append(rows)
        print(tag, rows.shape, 'ko', int(rows[:, 11913 + 4].sum()),
              'ladder', int(rows[:, 2 * 361:3 * 361].sum() + rows[:, 15 * 361:16 * 361].sum()))

    import torch
    from maru_b200 import netspec
    net = netspec.build(**netspec.CONFIGS['b2c64'], seed=0)
    path = '/tmp/golden_b2c64.model'
    netspec.export(net, path)
    digest = hashlib.sha256(b''.join(v.detach().numpy().tobytes()
                                     for v in net.state_dict().values())).hexdigest()
    proc = ref.RefProcessor(path, gpus=[-1], deterministic=True)
    outs = {}
    for (tag, _, _, _), rows in zip(SETS, all_rows):
        outs['out_' + tag] = proc.execute(rows)
    np.savez_compressed(HERE / 'outputs_b2c64.npz', weights_sha256=digest,
                        torch_version=torch.__version__, **outs)
    print('weights sha256', digest)


if __name__ == '__main__':
    main()
