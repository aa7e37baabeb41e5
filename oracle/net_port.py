"""TEST INFRASTRUCTURE ONLY (oracle "port" of the network). Never imported by the product.

Evaluates a `.b200` weight pack in plain PyTorch fp32 on the CPU, stage by stage, in the SAME order
as the engine's launch sequence (maru_b200/csrc/engine.cu: launch_trunk), so that GPU tests can
compare every intermediate activation buffer and localise a wrong kernel. It restates the builder-
defined architecture of maru_b200/netspec.py with BN already folded (maru_b200/convert.py).

Pinned: tests/test_convert.py checks that this port reproduces the reference runtime
(`deepgo::Model::forward` on the exported TorchScript file, via oracle/_ref) to fp32 round-off when
`emulate_bf16=False` and weights are left unrounded; that pins convert.py's folding as well.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

CELLS = 361
STEM_CIN = 64


def read_pack(path: str) -> Tuple[List[int], Dict[str, torch.Tensor]]:
    raw = open(path, 'rb').read()
    magic, version, n = struct.unpack_from('<8sII', raw, 0)
    assert magic == b'MARUB200' and version == 1
    arch = list(struct.unpack_from('<8i', raw, 16))
    tensors = {}
    for i in range(n):
        name, dtype, ndim, d0, d1, d2, d3, off, nbytes = struct.unpack_from('<48sII4IQQ', raw, 48 + i * 88)
        name = name.split(b'\0')[0].decode()
        dims = [d0, d1, d2, d3][:ndim]
        buf = raw[off:off + nbytes]
        if dtype == 1:
            u16 = np.frombuffer(buf, np.uint16).astype(np.uint32) << 16
            t = torch.from_numpy(u16.view(np.float32).copy())
        else:
            t = torch.from_numpy(np.frombuffer(buf, np.float32).copy())
        tensors[name] = t.reshape(dims)
    return arch, tensors


def bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def trunk_input(rows: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """[n,11920] reference rows -> (x0 [n,64,19,19], mask [n,1,19,19]) as encode.cu builds them."""
    n = rows.shape[0]
    planes = rows[:, :33 * CELLS].reshape(n, 33, 19, 19)
    info = rows[:, 33 * CELLS:]
    m = planes[:, 32:33]
    x0 = torch.zeros(n, STEM_CIN, 19, 19)
    x0[:, :33] = (planes != 0).float()
    inf = bf16(info)
    hi = bf16(info[:, 2])
    lo = bf16(info[:, 2] - hi)
    inf[:, 2] = hi
    x0[:, 33:40] = inf[:, :, None, None] * m
    x0[:, 40] = lo[:, None, None] * m[:, 0]
    return x0, m


def _conv(x, w, b, taps):
    """w: [taps][cout][cin] -> conv2d; returns fp32."""
    t, cout, cin = w.shape
    if taps == 9:
        k = w.reshape(3, 3, cout, cin).permute(2, 3, 0, 1)
        return F.conv2d(x, k, b, padding=1)
    return F.conv2d(x, w.reshape(cout, cin, 1, 1), b)


def staged_forward(arch: List[int], T: Dict[str, torch.Tensor], rows: np.ndarray,
                   emulate_bf16: bool = True):
    """Returns (stages, out [n,2169], logits [n,2,361]). stages: list of (name, buffer id, tensor
    [n,C,19,19]) in launch order; buffer ids follow MARU_BUF_* (1 h, 2 r, 3 s, 4 qkv, 5 o, 6 g)."""
    blocks, C, attn_every, head_dim, Ch, Cv = arch[:6]
    rnd = bf16 if emulate_bf16 else (lambda v: v)
    rows_t = torch.from_numpy(np.ascontiguousarray(rows, np.float32))
    n = rows_t.shape[0]
    x0, m = trunk_input(rows_t)
    stages = []
    with torch.no_grad():
        h = rnd(torch.relu(_conv(x0, T['stem.w'], T['stem.b'], 9)) * m)
        stages.append(('stem', 1, h))
        ai = 0
        for i in range(blocks):
            b = f'blk{i}'
            r = rnd(torch.relu(_conv(h, T[b + '.reduce.w'], T[b + '.reduce.b'], 1)) * m)
            stages.append((b + '.reduce', 2, r))
            for j in range(2):
                s = rnd(torch.relu(_conv(r, T[f'{b}.in{j}.c1.w'], T[f'{b}.in{j}.c1.b'], 9)) * m)
                stages.append((f'{b}.in{j}.c1', 3, s))
                r = rnd(torch.relu(_conv(s, T[f'{b}.in{j}.c2.w'], T[f'{b}.in{j}.c2.b'], 9) + r) * m)
                stages.append((f'{b}.in{j}.c2', 2, r))
            h = rnd(torch.relu(_conv(r, T[b + '.expand.w'], T[b + '.expand.b'], 1) + h) * m)
            stages.append((b + '.expand', 1, h))
            if attn_every > 0 and (i + 1) % attn_every == 0 and f'att{ai}.qkv.w' in T:
                a = f'att{ai}'
                ai += 1
                qkv = rnd(_conv(h, T[a + '.qkv.w'], T[a + '.qkv.b'], 1) * m)
                stages.append((a + '.qkv', 4, qkv))
                H = C // head_dim
                q = qkv[:, :C].reshape(n, H, head_dim, CELLS).transpose(-1, -2)
                k = qkv[:, C:2 * C].reshape(n, H, head_dim, CELLS)
                v = qkv[:, 2 * C:].reshape(n, H, head_dim, CELLS).transpose(-1, -2)
                sc = torch.matmul(q, k)                      # already in log2 units (q pre-scaled)
                sc = sc.masked_fill(m.reshape(n, 1, 1, CELLS) < 0.5, float('-inf'))
                p = torch.softmax(sc * float(np.log(2.0)), dim=-1)
                o = torch.matmul(p, v).transpose(-1, -2).reshape(n, C, 19, 19)
                o = rnd(o * m)
                stages.append((a + '.core', 5, o))
                h = rnd((_conv(o, T[a + '.proj.w'], T[a + '.proj.b'], 1) + h) * m)
                stages.append((a + '.proj', 1, h))
        g = rnd(torch.relu(_conv(h, T['head.g.w'], T['head.g.b'], 1)) * m)
        stages.append(('head.g', 6, g))
        z = F.conv2d(g, T['head.planes.w'].reshape(6, Ch, 1, 1), T['head.planes.b']).reshape(n, 6, CELLS)
        mf = m.reshape(n, 1, CELLS)
        pol = torch.softmax(z[:, 0:2].masked_fill(mf < 0.5, float('-inf')), dim=-1)
        ter = torch.softmax(z[:, 2:5], dim=1) * mf
        val = torch.sigmoid(z[:, 5:6]) * mf
        gf = g.reshape(n, Ch, CELLS)
        pooled = torch.cat([gf.sum(-1) / mf.sum(-1), gf.max(-1)[0]], dim=1)
        hid = torch.relu(F.linear(pooled, T['head.fc1.w'], T['head.fc1.b']))
        v3 = F.linear(hid, T['head.fc2.w'], T['head.fc2.b'])
        scal = torch.cat([torch.sigmoid(v3[:, 0:1]), v3[:, 1:3]], dim=1)
        out = torch.cat([torch.cat([pol, ter, val], dim=1).reshape(n, 6 * CELLS), scal], dim=1)
    return stages, out.numpy(), z[:, 0:2].numpy()
