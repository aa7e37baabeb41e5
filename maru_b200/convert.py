"""The Python loader: TorchScript `.model` -> `.b200` weight pack for libmaru_b200.so.

This is the only place PyTorch touches the product: `torch.jit.load(path).state_dict()` (what the
reference does in C++ at src/deepgo/native/cpp/Model.cpp:77) -> BatchNorm folding -> bf16 cast ->
a flat binary pack that `maru_eval_create` maps into device buffers (maru_b200/csrc/engine.cu).
No forward pass is ever run here.

usage:  python -m maru_b200.convert <file>.model [<out>.b200]

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


def convert(model_path: str, out_path: str | None = None) -> str:
    """`.model` (TorchScript archive) -> `<model_path>.b200`. Returns the pack path."""
    mod = torch.jit.load(model_path, map_location='cpu')
    sd = {k: v.detach().cpu() for k, v in mod.state_dict().items()}
    if 'arch' not in sd:
        raise ValueError(f'{model_path}: no `arch` buffer - not a netspec.MaruNet export; '
                         'released Maru weights need a name map (SURVEY f4)')
    arch, tensors = fold_state_dict(sd)
    out_path = out_path or (model_path + '.b200')
    write_pack(arch, tensors, out_path)
    return out_path


def convert_state_dict(sd: Dict[str, torch.Tensor], out_path: str) -> str:
    arch, tensors = fold_state_dict({k: v.detach().cpu() for k, v in sd.items()})
    write_pack(arch, tensors, out_path)
    return out_path


if __name__ == '__main__':
    if len(sys.argv) < 2:
        print(__doc__)
        sys.exit(2)
    print(convert(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None))
