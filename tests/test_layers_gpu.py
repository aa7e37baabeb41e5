"""GPU: every activation buffer of the trunk, stage by stage, against the torch port of the packed
network (oracle/net_port.py, bf16 storage emulated). Localises a wrong kernel; on a stem mismatch
it also probes the UMMA descriptor variants so one GPU run tells which encoding is right."""
import numpy as np
import pytest

from conftest import load_positions

pytestmark = pytest.mark.gpu

BUF_CH = {1: 'C', 2: 'C/2', 3: 'C/2', 4: '3C', 5: 'C', 6: 'Ch'}


def run_stages(ev, arch, T, rows, upto=None):
    from oracle import net_port
    C, Ch = arch[1], arch[4]
    chans = {1: C, 2: C // 2, 3: C // 2, 4: 3 * C, 5: C, 6: Ch}
    stages, out, logits = net_port.staged_forward(arch, T, rows, emulate_bf16=True)
    report = []
    n = len(rows)
    for k, (name, buf, want) in enumerate(stages):
        if upto is not None and k >= upto:
            break
        ev.debug_set(0, k + 1)
        ev.forward(rows)
        got = ev.debug_planes(buf, n, chans[buf])
        want = want.reshape(n, chans[buf], 361).numpy()
        err = np.abs(got - want).max()
        scale = np.abs(want).max() + 1e-6
        report.append((name, float(err), float(scale)))
    ev.debug_set(0, -1)
    return report, out, logits


@pytest.mark.parametrize('name', ['b2c64', 'b4c128', 'b10c256'])
def test_every_stage_matches_port(built_lib, model_dir, name):
    import maru_b200
    from oracle import net_port
    path = model_dir(name)
    arch, T = net_port.read_pack(str(path) + '.b200')
    rows = np.concatenate([load_positions('19')[0][:5], load_positions('9')[0][:3],
                           load_positions('9x13')[0][24:27]])
    ev = maru_b200.Evaluator(path, gpu=0, max_batch=16, flags=1)
    try:
        report, out, logits = run_stages(ev, arch, T, rows, upto=1)
        if report[0][1] > 0.05 * report[0][2]:
            # bring-up aid: which LBO/SBO reading of the SWIZZLE_NONE descriptor does the HW use?
            probe = {}
            for bits in (1, 2, 3):
                ev.debug_set(1, bits)
                probe[bits] = run_stages(ev, arch, T, rows, upto=1)[0][0][1]
            ev.debug_set(1, 0)
            pytest.fail(f'stem mismatch err={report[0]} ; descriptor probe (bit0 swap A, bit1 swap B): {probe}')
        report, out, logits = run_stages(ev, arch, T, rows)
        core = [r for r in report if r[0].endswith('.core')]
        if core and core[0][1] > 0.05 * core[0][2] + 0.02:
            # bring-up aid for the tcgen05 attention: which reading of the MN-major V descriptor?
            probe = {}
            for bits in (4,):
                ev.debug_set(1, bits)
                rep2 = run_stages(ev, arch, T, rows)[0]
                probe[bits] = [r for r in rep2 if r[0].endswith('.core')][0]
            ev.debug_set(1, 0)
            pytest.fail(f'attention core mismatch {core[0]}; probe (4: legacy mma.sync kernel): {probe}')
        lines = [f'{n:18s} err={e:.4f} scale={s:.3f}' for n, e, s in report]
        print('\n'.join(lines))
        bad = [(n, e, s) for n, e, s in report if e > 0.02 * s + 0.02]
        assert not bad, 'stage mismatch: ' + str(bad) + '\n' + '\n'.join(lines)
        got = ev.forward(rows)
        glog = ev.debug_policy_logits(len(rows))
        mask = rows[:, 32 * 361:33 * 361] > 0
        assert np.abs(glog - logits)[np.repeat(mask[:, None], 2, 1)].max() < 2e-2
        assert np.abs(got - out).max() < 1e-2
    finally:
        ev.close()


def test_flow_kernel_matches_per_layer_launches(built_lib, model_dir):
    """The persistent multi-layer conv kernel (conv_flow.cu, the default) against one PDL launch per layer
    (debug bit 5): same MMAs in the same order, but the residual is added by the epilogue in fp32 instead
    of an identity MMA, so residual layers may differ by one bf16 rounding. Every activation buffer left
    behind by a whole forward and the outputs must agree to bf16 noise, the flow path must be
    run-to-run deterministic, and it must really have replaced the per-layer launches."""
    import maru_b200
    from maru_b200 import synth
    for name, sizes in (('b2c64', (1, 13)), ('b4c128', (1, 2, 40, 129, 300, 1100)), ('b10c256', (7,))):
        ev = maru_b200.Evaluator(model_dir(name), gpu=0, max_batch=1100 if name == 'b4c128' else 64, flags=maru_b200.capi.FLAG_NO_GRAPH)
        try:
            C = ev.info().channels
            chans = {1: C, 2: C // 2, 3: C // 2, 4: 3 * C, 5: C, 6: 32}
            for n in sizes:
                st = synth.random_states(n, 19 if n != 13 else 9, seed=5 + n)
                ev.debug_set(1, 32)
                want = ev.forward_states(st)
                launches = ev.info().launches_per_forward
                wbuf = {b: ev.debug_planes(b, n, c) for b, c in chans.items()}
                ev.debug_set(1, 128)       # the flow kernels at every batch size (default: for 249..354 positions on 148 SMs)
                got = ev.forward_states(st)
                if name != 'b10c256':      # C=256: the 3x3 weights (295 KB) and the 64 KB 1x1 slabs do not fit a chain
                    assert ev.info().launches_per_forward < launches, (name, n)    # layers really were chained
                gbuf = {b: ev.debug_planes(b, n, c) for b, c in chans.items()}
                again = ev.forward_states(st)
                assert np.array_equal(got, again), (name, n)
                for b in chans:
                    scale = np.abs(wbuf[b]).max() + 1e-6
                    err = np.abs(gbuf[b] - wbuf[b]).max()
                    assert err <= 0.02 * scale + 1e-3, (name, n, b, float(err), float(scale))
                assert np.abs(got - want).max() < 2e-3, (name, n, float(np.abs(got - want).max()))
        finally:
            ev.close()


def test_attention_with_and_without_the_row_shift(built_lib, model_dir):
    """attention.cu: a warp whose rows' scores are bounded by 64 (|q_i| max_j |k_j|, log2 units) skips the subtraction
    of the row shift c_i before the exponential; debug bit 23 forces the shifted path everywhere. Softmax is shift
    invariant: the attention output (buffer 5) and the network outputs must agree to the rounding of P (bf16), on the
    19x19 frame and on a compact frame."""
    import maru_b200
    from maru_b200 import synth
    for name in ('b2c64', 'b4c128'):
        ev = maru_b200.Evaluator(model_dir(name), gpu=0, max_batch=64, flags=maru_b200.capi.FLAG_NO_GRAPH)
        try:
            C = ev.info().channels
            for size, n in ((19, 33), (9, 64)):
                st = synth.random_states(n, size, seed=70 + size)
                ev.debug_set(1, 32)
                got = ev.forward_states(st)
                gbuf = ev.debug_planes(5, n, C)
                ev.debug_set(1, 32 | (1 << 23))
                want = ev.forward_states(st)
                wbuf = ev.debug_planes(5, n, C)
                scale = np.abs(wbuf).max() + 1e-6
                assert np.abs(gbuf - wbuf).max() <= 0.02 * scale + 1e-3, (name, size)
                assert np.abs(got - want).max() < 2e-3, (name, size, float(np.abs(got - want).max()))
                assert not np.array_equal(gbuf, wbuf), (name, size)      # the unshifted path really ran by default
        finally:
            ev.close()


def test_ksplit_kernel_matches_conv_gemm(built_lib, model_dir):
    """C = 256 nets: the CTA-pair kernel conv_pair.cu (3x3 128 -> 128, cta_group::2), its fallback conv_ksplit.cu (half-slab
    pipeline) and the 256-channel-per-CTA projections of conv_wide.cu, all against conv_gemm (debug bit 4). Same MMAs, but channels 0-63 of all taps are accumulated before 64-127:
    equal to fp32 rounding of the accumulator, i.e. at most a bf16 ulp per layer; deterministic run to run."""
    import maru_b200
    from maru_b200 import synth
    ev = maru_b200.Evaluator(model_dir('b10c256'), gpu=0, max_batch=64, flags=maru_b200.capi.FLAG_NO_GRAPH)
    try:
        for n in (1, 37):
            st = synth.random_states(n, 19, seed=40 + n)
            ev.debug_set(1, 16)
            want = ev.forward_states(st)
            wbuf = ev.debug_planes(2, n, 128)
            ev.debug_set(1, 1 << 20)           # conv_ksplit for the 128 -> 128 layers (the fallback of conv_pair)
            mid = ev.forward_states(st)
            assert np.abs(mid - want).max() < 5e-3, float(np.abs(mid - want).max())
            ev.debug_set(1, 0)                 # default: conv_pair (CTA pairs) + conv_wide
            got = ev.forward_states(st)
            gbuf = ev.debug_planes(2, n, 128)
            assert np.array_equal(got, ev.forward_states(st))
            # two bf16 pipelines that round differently in 45 of the 79 layers: a few bf16 ulps (0.4-0.8 % each) on
            # the largest activations after ten blocks, far inside the gates against the reference
            scale = np.abs(wbuf).max() + 1e-6
            err = float(np.abs(gbuf - wbuf).max())
            assert err <= 0.05 * scale + 1e-3, (err, float(scale))
            assert np.abs(got - want).max() < 5e-3, float(np.abs(got - want).max())
    finally:
        ev.close()


@pytest.mark.parametrize('size', [3, 5, 7])
def test_small_boards_skip_key_parts(built_lib, model_dir, size):
    """Attention skips key parts (80 padded rows) without an on-board cell: a 3x3 board lives in ONE of the five parts,
    5x5 in two, 7x7 in three — the short step sequences (and their epilogue placement) against the torch port."""
    import maru_b200
    from maru_b200 import synth
    from oracle import net_port
    path = model_dir('b4c128')
    arch, T = net_port.read_pack(str(path) + '.b200')
    ev = maru_b200.Evaluator(path, gpu=0, max_batch=16)
    try:
        st = synth.random_states(9, size, seed=70 + size)
        rows = ev.encode_planes(st)
        _, want, logits = net_port.staged_forward(arch, T, rows, emulate_bf16=True)
        got = ev.forward_states(st)
        assert np.array_equal(got, ev.forward_states(st))
        mask = rows[:, 32 * 361:33 * 361] > 0
        assert mask.sum(1).max() == size * size
        glog = ev.debug_policy_logits(len(rows))
        logits = np.asarray(logits)
        want = np.asarray(want)
        assert np.abs(glog - logits)[np.repeat(mask[:, None], 2, 1)].max() < 2e-2
        assert np.abs(got - want).max() < 1e-2
    finally:
        ev.close()
