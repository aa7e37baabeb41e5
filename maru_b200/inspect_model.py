"""SURVEY f4 — introspect a TorchScript `.model` archive and print a name-map template for the loader.

The reference loads its network opaquely (`torch::jit::load`, src/deepgo/native/cpp/Model.cpp:77): the released
Maru `.model` files (README.md:10) ship no Python source, and none is available offline. This tool lists what
such an archive contains — parameters / buffers with shapes, convolution kernel sizes, channel histogram, the
scripted `forward` source when present — and writes a JSON name map skeleton

    {"arch": [blocks, channels, attn_every, head_dim, head_channels, value_channels],
     "rename": {"<archive key>": "<netspec key or empty to drop>", ...}}

that `python -m maru_b200.convert <file>.model --name-map=<map>.json` applies before BatchNorm folding. Keys
that already carry a netspec.MaruNet name map to themselves; everything else is left for the maintainer to fill
in against `maru_b200.convert.expected_keys(arch)` (printed with their shapes). No forward pass is run.

usage:  python -m maru_b200.inspect_model <file>.model [<map-template>.json]
"""
from __future__ import annotations

import collections
import json
import sys
from typing import Dict

import torch

from . import convert


def describe(sd: Dict[str, torch.Tensor]) -> dict:
    """Structure summary of a state_dict: per-tensor shapes plus the statistics a maintainer needs to guess
    the architecture (number of 3x3 / 1x1 convolutions, channel widths, BatchNorm count)."""
    tensors = {k: [int(d) for d in v.shape] for k, v in sd.items()}
    conv3 = [k for k, s in tensors.items() if len(s) == 4 and s[2:] == [3, 3]]
    conv1 = [k for k, s in tensors.items() if len(s) == 4 and s[2:] == [1, 1]]
    linear = [k for k, s in tensors.items() if len(s) == 2]
    bn = sorted({k.rsplit('.', 1)[0] for k in tensors if k.endswith('.running_var')})
    widths = collections.Counter(s[0] for k, s in tensors.items() if len(s) == 4)
    return dict(n_tensors=len(tensors), n_params=int(sum(v.numel() for v in sd.values())), conv3x3=conv3,
                conv1x1=conv1, linear=linear, batchnorm=bn,
                out_channel_histogram={str(k): v for k, v in sorted(widths.items())}, tensors=tensors)


def guess_arch(sd: Dict[str, torch.Tensor]) -> list | None:
    """The `arch` buffer when present, else a guess from the tensor shapes assuming the netspec family
    (stem 3x3 from 33 planes, four C/2 3x3 convolutions per block)."""
    if 'arch' in sd:
        return [int(v) for v in sd['arch'].tolist()][:6]
    stem = [v for v in sd.values() if v.dim() == 4 and tuple(v.shape[1:]) == (33, 3, 3)]
    if not stem:
        return None
    C = int(stem[0].shape[0])
    inner = [v for v in sd.values() if v.dim() == 4 and tuple(v.shape) == (C // 2, C // 2, 3, 3)]
    qkv = [v for v in sd.values() if v.dim() >= 2 and int(v.shape[0]) == 3 * C and int(v.shape[1]) == C]
    blocks = len(inner) // 4
    n_attn = len(qkv)
    return [blocks, C, (blocks // n_attn) if n_attn else 0, 32, 32, 64]


def template(sd: Dict[str, torch.Tensor]) -> dict:
    arch = guess_arch(sd)
    want = convert.expected_keys(arch) if arch else {}
    rename = {}
    for k in sd:
        if k == 'arch':
            continue
        if k in want:
            rename[k] = k
        elif k.endswith('num_batches_tracked'):
            rename[k] = ''
        else:
            rename[k] = None                  # to be filled in by the maintainer
    return dict(arch=arch, rename=rename,
                unmatched_netspec_keys={k: list(s) for k, s in want.items() if k not in sd})


def inspect(path: str) -> dict:
    mod = torch.jit.load(path, map_location='cpu')
    sd = {k: v.detach().cpu() for k, v in mod.state_dict().items()}
    info = describe(sd)
    try:
        info['forward_source'] = mod.code
    except Exception as exc:                  # archives scripted without source still load
        info['forward_source'] = f'<unavailable: {exc}>'
    info['template'] = template(sd)
    return info


def main(argv) -> int:
    if len(argv) < 2:
        print(__doc__)
        return 2
    info = inspect(argv[1])
    t = info.pop('template')
    src = info.pop('forward_source')
    tensors = info.pop('tensors')
    print(json.dumps(info, indent=1))
    print(f'--- {len(tensors)} tensors')
    for k, s in tensors.items():
        print(f'  {k:56s} {s}')
    print('--- forward source\n' + src[:4000])
    done = sum(1 for v in t['rename'].values() if v)
    print(f'--- name map: {done} of {len(t["rename"])} keys matched, {len(t["unmatched_netspec_keys"])} netspec keys open; '
          f'guessed arch {t["arch"]}')
    if len(argv) > 2:
        with open(argv[2], 'w') as fh:
            json.dump(dict(arch=t['arch'], rename=t['rename']), fh, indent=1)
        print('wrote', argv[2])
    return 0


if __name__ == '__main__':
    sys.exit(main(sys.argv))
