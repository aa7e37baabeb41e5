"""The Python loader: TorchScript `.model` -> `.b200` weight pack for libmaru_b200.so.

This is the only place PyTorch touches the product: `torch.jit.load(path).state_dict()` (what the
reference does in C++ at src/deepgo/native/cpp/Model.cpp:77) -> BatchNorm folding -> bf16 cast ->
a flat binary pack that `maru_eval_create` maps into device buffers (maru_b200/csrc/engine.cu).
No forward pass is ever run here.

usage:  python -m maru_b200.convert <file>.model [<out>.b200] [--name-map=<map>.json]

Pack layout (little endian):
  header  : magic "MARUB200", u32 version=1, u32 n_tensors, i32 arch[8]
            arch = blocks, channels, attn_every, head_dim, head_channels, value_channels, 0, 0
  entries : n_tensors x { char name[48]; u32 dtype (0 f32, 1 bf16); u32 ndim; u32 dims[4];
                          u64 offset; u64 nbytes }
  data    : 64-byte aligned blobs
Conv weights are stored [tap][cout][cin] bf16 with BN scale folded in; biases fp32.
"""
from __future__ import annotations

import math
import struct
import sys
from typing import Dict, Tuple

import numpy as np
import torch

BN_EPS = 1e-5
STEM_CIN = 64
STEM_INFO0 = 33
STEM_KOMI_LO = 40
LOG2E = 1.4426950408889634


def _bn_fold(sd: Dict[str, torch.Tensor], prefix: str) -> Tuple[torch.Tensor, torch.Tensor]:
    w, b = sd[prefix + '.weight'].double(), sd[prefix + '.bias'].double()
    mu, var = sd[prefix + '.running_mean'].double(), sd[prefix + '.running_var'].double()
    scale = w / torch.sqrt(var + BN_EPS)
    return scale, b - mu * scale


def _taps(w: torch.Tensor) -> torch.Tensor:
    """[cout, cin, kh, kw] -> [kh*kw, cout, cin]; tap t = ky*kw + kx (kernel: dy=t/3-1, dx=t%3-1)."""
    cout, cin, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin).contiguous()


def fold_state_dict(sd: Dict[str, torch.Tensor]) -> Tuple[list, Dict[str, torch.Tensor]]:
    """Returns (arch, tensors) where tensors maps pack names to fp64/fp32 torch tensors
    (conv weights [tap][cout][cin] pre-bf16, biases)."""
    arch = [int(v) for v in sd['arch'].tolist()]
    blocks, C, attn_every, head_dim, Ch, Cv = arch[:6]
    n_attn = blocks // attn_every if attn_every > 0 else 0
    out: Dict[str, torch.Tensor] = {}

    # stem: conv3x3(33->C) + linear(info 7->C), BN folded; info enters as centre-tap-only channels
    scale, shift = _bn_fold(sd, 'stem_bn')
    wc = sd['stem_conv.weight'].double() * scale[:, None, None, None]
    wi = sd['stem_info.weight'].double() * scale[:, None]            # [C, 7]
    w = torch.zeros(9, C, STEM_CIN, dtype=torch.float64)
    w[:, :, :33] = _taps(wc)
    w[4, :, STEM_INFO0:STEM_INFO0 + 7] = wi
    w[4, :, STEM_KOMI_LO] = wi[:, 2]                                  # low half of the komi split
    out['stem.w'], out['stem.b'] = w, shift

    def convbn(dst: str, src: str):
        s, b = _bn_fold(sd, src + '.bn')
        out[dst + '.w'] = _taps(sd[src + '.conv.weight'].double() * s[:, None, None, None])
        out[dst + '.b'] = b

    for i in range(blocks):
        convbn(f'blk{i}.reduce', f'blocks.{i}.reduce')
        for j in range(2):
            convbn(f'blk{i}.in{j}.c1', f'blocks.{i}.inner{j}.c1')
            convbn(f'blk{i}.in{j}.c2', f'blocks.{i}.inner{j}.c2')
        convbn(f'blk{i}.expand', f'blocks.{i}.expand')

    for k in range(n_attn):
        s, b = _bn_fold(sd, f'attn.{k}.norm')
        wq = sd[f'attn.{k}.qkv.weight'].double().reshape(3 * C, C)
        bq = sd[f'attn.{k}.qkv.bias'].double()
        w = wq * s[None, :]
        bias = wq @ b + bq
        qs = LOG2E / math.sqrt(head_dim)                              # softmax becomes a bare exp2
        w[:C] *= qs
        bias[:C] *= qs
        out[f'att{k}.qkv.w'], out[f'att{k}.qkv.b'] = w.reshape(1, 3 * C, C), bias
        out[f'att{k}.proj.w'] = sd[f'attn.{k}.proj.weight'].double().reshape(1, C, C)
        out[f'att{k}.proj.b'] = sd[f'attn.{k}.proj.bias'].double()

    s, b = _bn_fold(sd, 'heads.g.bn')
    out['head.g.w'] = _taps(sd['heads.g.conv.weight'].double() * s[:, None, None, None])
    out['head.g.b'] = b
    out['head.planes.w'] = sd['heads.planes.weight'].double().reshape(6, Ch)
    out['head.planes.b'] = sd['heads.planes.bias'].double()
    out['head.fc1.w'] = sd['heads.fc1.weight'].double()
    out['head.fc1.b'] = sd['heads.fc1.bias'].double()
    out['head.fc2.w'] = sd['heads.fc2.weight'].double()
    out['head.fc2.b'] = sd['heads.fc2.bias'].double()
    return arch, out


def _is_conv_weight(name: str) -> bool:
    return name.endswith('.w') and not name.startswith('head.planes') and not name.startswith('head.fc')


def write_pack(arch: list, tensors: Dict[str, torch.Tensor], path: str) -> None:
    entries = []
    blobs = []
    for name, t in tensors.items():
        if _is_conv_weight(name):
            arr = t.float().to(torch.bfloat16).contiguous().view(torch.int16).numpy()
            dtype = 1
        else:
            arr = t.float().contiguous().numpy()
            dtype = 0
        entries.append((name, dtype, list(t.shape), arr.tobytes()))
    table = 48 + len(entries) * 88
    off = (table + 63) // 64 * 64
    recs = b''
    for name, dtype, shape, raw in entries:
        dims = (shape + [1, 1, 1, 1])[:4]
        assert len(name) < 48 and len(shape) <= 4
        recs += struct.pack('<48sII4IQQ', name.encode(), dtype, len(shape), *dims, off, len(raw))
        blobs.append((off, raw))
        off = (off + len(raw) + 63) // 64 * 64
    with open(path, 'wb') as f:
        f.write(struct.pack('<8sII8i', b'MARUB200', 1, len(entries), *(arch + [0] * 8)[:8]))
        f.write(recs)
        for o, raw in blobs:
            f.seek(o)
            f.write(raw)
        f.truncate(off)


def apply_name_map(sd: Dict[str, torch.Tensor], name_map: dict) -> Dict[str, torch.Tensor]:
    """a TorchScript archive whose parameters are named differently from netspec.MaruNet
    (a released Maru `.model`; the reference loads it opaquely, Model.cpp:77) is converted through a name
    map, a JSON object written from the template `python -m maru_b200.inspect_model <file>.model` prints:

      {"arch": [blocks, channels, attn_every, head_dim, head_channels, value_channels],
       "rename": {"<their key>": "<netspec key>", ...}}

    Every netspec key must be produced exactly once; tensors are checked against the shapes `arch` implies
    by fold_state_dict (a wrong map fails loudly, it never produces a silently different network)."""
    rename = name_map.get('rename', {})
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        nk = rename.get(k, k)
        if nk is None or nk == '':
            continue                      # explicitly dropped (e.g. num_batches_tracked)
        if nk in out:
            raise ValueError(f'name map sends two tensors to {nk!r}')
        out[nk] = v
    if 'arch' in name_map:
        a = [int(x) for x in name_map['arch']]
        out['arch'] = torch.tensor(a + [0] * (8 - len(a)), dtype=torch.int32)
    return out


def expected_keys(arch: list) -> Dict[str, tuple]:
    """netspec.MaruNet state_dict keys (those the loader reads) -> shapes implied by `arch`."""
    blocks, C, attn_every, head_dim, Ch, Cv = [int(v) for v in arch[:6]]
    n_attn = blocks // attn_every if attn_every > 0 else 0
    keys: Dict[str, tuple] = {}

    def bn(prefix, c):
        for f in ('weight', 'bias', 'running_mean', 'running_var'):
            keys[f'{prefix}.{f}'] = (c,)

    def convbn(prefix, cout, cin, k):
        keys[f'{prefix}.conv.weight'] = (cout, cin, k, k)
        bn(f'{prefix}.bn', cout)

    keys['stem_conv.weight'] = (C, 33, 3, 3)
    keys['stem_info.weight'] = (C, 7)
    bn('stem_bn', C)
    for i in range(blocks):
        convbn(f'blocks.{i}.reduce', C // 2, C, 1)
        for j in range(2):
            convbn(f'blocks.{i}.inner{j}.c1', C // 2, C // 2, 3)
            convbn(f'blocks.{i}.inner{j}.c2', C // 2, C // 2, 3)
        convbn(f'blocks.{i}.expand', C, C // 2, 1)
    for k in range(n_attn):
        bn(f'attn.{k}.norm', C)
        keys[f'attn.{k}.qkv.weight'] = (3 * C, C, 1, 1)
        keys[f'attn.{k}.qkv.bias'] = (3 * C,)
        keys[f'attn.{k}.proj.weight'] = (C, C, 1, 1)
        keys[f'attn.{k}.proj.bias'] = (C,)
    convbn('heads.g', Ch, C, 1)
    keys['heads.planes.weight'] = (6, Ch, 1, 1)
    keys['heads.planes.bias'] = (6,)
    keys['heads.fc1.weight'] = (Cv, 2 * Ch)
    keys['heads.fc1.bias'] = (Cv,)
    keys['heads.fc2.weight'] = (3, Cv)
    keys['heads.fc2.bias'] = (3,)
    return keys


def check_state_dict(sd: Dict[str, torch.Tensor]) -> None:
    """Raises ValueError listing every missing key / wrong shape (numel must match; conv weights may be
    stored with or without the trailing 1x1 dimensions)."""
    if 'arch' not in sd:
        raise ValueError('no `arch` buffer')
    problems = []
    for k, shape in expected_keys([int(v) for v in sd['arch'].tolist()]).items():
        if k not in sd:
            problems.append(f'missing {k} {shape}')
        elif int(sd[k].numel()) != int(np.prod(shape)):
            problems.append(f'{k}: shape {tuple(sd[k].shape)}, expected {shape}')
    if problems:
        raise ValueError('state_dict does not match the architecture in `arch`:\n  ' + '\n  '.join(problems))


def convert(model_path: str, out_path: str | None = None, name_map: dict | None = None) -> str:
    """`.model` (TorchScript archive) -> `<model_path>.b200`. Returns the pack path."""
    mod = torch.jit.load(model_path, map_location='cpu')
    sd = {k: v.detach().cpu() for k, v in mod.state_dict().items()}
    if name_map is not None:
        sd = apply_name_map(sd, name_map)
    if 'arch' not in sd:
        raise ValueError(f'{model_path}: no `arch` buffer - not a netspec.MaruNet export; '
                         'released Maru weights need a name map (SURVEY f4): '
                         'python -m maru_b200.inspect_model <file>.model prints a template')
    check_state_dict(sd)
    arch, tensors = fold_state_dict(sd)
    out_path = out_path or (model_path + '.b200')
    write_pack(arch, tensors, out_path)
    return out_path


def convert_state_dict(sd: Dict[str, torch.Tensor], out_path: str, name_map: dict | None = None) -> str:
    sd = {k: v.detach().cpu() for k, v in sd.items()}
    if name_map is not None:
        sd = apply_name_map(sd, name_map)
    check_state_dict(sd)
    arch, tensors = fold_state_dict(sd)
    write_pack(arch, tensors, out_path)
    return out_path


if __name__ == '__main__':
    if len(sys.argv) < 2:
        print(__doc__)
        sys.exit(2)
    args = [a for a in sys.argv[1:] if not a.startswith('--name-map=')]
    maps = [a.split('=', 1)[1] for a in sys.argv[1:] if a.startswith('--name-map=')]
    nm = None
    if maps:
        import json
        with open(maps[0]) as fh:
            nm = json.load(fh)
    print(convert(args[0], args[1] if len(args) > 1 else None, nm))
