"""CPU: the loader (maru_b200/convert.py: BN fold, q pre-scale, stem info channels, bf16 cast) and
the torch port of the packed network (oracle/net_port.py) against the reference runtime."""
import numpy as np
import pytest

from conftest import GOLDEN, load_positions


def centered_log_policy(out, rows):
    mask = rows[:, 32 * 361:33 * 361] > 0
    lp = np.log(np.maximum(out[:, :361], 1e-30))
    lp = np.where(mask, lp, 0.0)
    lp -= (lp.sum(1) / mask.sum(1))[:, None]
    return np.where(mask, lp, 0.0)


def test_pack_roundtrip_shapes(model_dir):
    from oracle import net_port
    arch, T = net_port.read_pack(str(model_dir('b2c64')) + '.b200')
    assert arch[:6] == [2, 64, 1, 32, 32, 64]
    assert T['stem.w'].shape == (9, 64, 64)
    assert T['blk0.in1.c2.w'].shape == (9, 32, 32)
    assert T['att1.qkv.w'].shape == (1, 192, 64)
    # info scalars enter only through the centre tap
    assert T['stem.w'][[0, 1, 2, 3, 5, 6, 7, 8]][:, :, 33:41].abs().sum() == 0
    assert T['stem.w'][4][:, 33:41].abs().sum() > 0
    assert T['stem.w'][:, :, 41:].abs().sum() == 0


def test_port_matches_golden_reference_outputs(model_dir):
    """golden outputs = deepgo::Processor::execute of the same seed-0 b2c64 export."""
    from oracle import net_port
    arch, T = net_port.read_pack(str(model_dir('b2c64')) + '.b200')
    gold = np.load(GOLDEN / 'outputs_b2c64.npz')
    for tag in ['9', '19', '9x13']:
        rows, _ = load_positions(tag)
        want = gold['out_' + tag]
        _, out, _ = net_port.staged_forward(arch, T, rows, emulate_bf16=False)
        assert np.abs(out - want).max() < 2e-3                       # weights rounded to bf16 only
        assert np.abs(centered_log_policy(out, rows) - centered_log_policy(want, rows)).max() < 1e-2
        assert np.abs(out[:, 2166] - want[:, 2166]).max() < 2e-3


def test_port_matches_live_reference(model_dir):
    from oracle import net_port, ref
    if not ref.available('nn'):
        pytest.skip('oracle/_ref nn not built')
    path = model_dir('b2c64', seed=3)
    arch, T = net_port.read_pack(str(path) + '.b200')
    rows, _, _ = ref.make_positions(12, 13, seed=9)
    want = ref.RefProcessor(str(path), gpus=[-1]).execute(rows)
    _, out, _ = net_port.staged_forward(arch, T, rows, emulate_bf16=True)
    assert np.abs(out - want).max() < 5e-3
    assert np.abs(out[:, 2166:] - want[:, 2166:]).max() < 5e-3


def test_name_map_converts_a_foreign_state_dict(tmp_path):
    """a state_dict whose keys are not netspec's (a released Maru archive) goes through a name map;
    the pack must be byte-identical to converting the netspec-named weights, and a wrong map must fail loudly."""
    import torch
    from maru_b200 import convert, inspect_model, netspec
    net = netspec.build(2, 64, 2, 32, 32, 64, seed=3)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    want = tmp_path / 'want.b200'
    convert.convert_state_dict(sd, str(want))
    # the same weights under foreign names, without the arch buffer
    foreign, rename = {}, {}
    for i, (k, v) in enumerate(sd.items()):
        if k == 'arch':
            continue
        fk = f'net.m{i:03d}.' + k.split('.')[-1]
        foreign[fk] = v
        rename[fk] = '' if k.endswith('num_batches_tracked') else k
    arch = [int(x) for x in sd['arch'].tolist()][:6]
    assert inspect_model.guess_arch(foreign) == arch           # shapes alone identify the netspec family
    t = inspect_model.template(foreign)
    assert len(t['unmatched_netspec_keys']) == len(convert.expected_keys(arch))
    got = tmp_path / 'got.b200'
    convert.convert_state_dict(foreign, str(got), name_map=dict(arch=arch, rename=rename))
    assert got.read_bytes() == want.read_bytes()
    # a map that loses a tensor, or swaps two of different shape, is rejected with the key named
    broken = dict(rename)
    lost = next(k for k, v in broken.items() if v == 'stem_conv.weight')
    broken[lost] = ''
    with pytest.raises(ValueError, match='stem_conv.weight'):
        convert.convert_state_dict(foreign, str(got), name_map=dict(arch=arch, rename=broken))
    swapped = dict(rename)
    a = next(k for k, v in swapped.items() if v == 'stem_conv.weight')
    b = next(k for k, v in swapped.items() if v == 'heads.fc2.bias')
    swapped[a], swapped[b] = swapped[b], swapped[a]
    with pytest.raises(ValueError, match='expected'):
        convert.convert_state_dict(foreign, str(got), name_map=dict(arch=arch, rename=swapped))


def test_inspect_model_reads_an_exported_archive(model_dir):
    from maru_b200 import inspect_model
    info = inspect_model.inspect(str(model_dir('b2c64')))
    assert info['template']['arch'][:2] == [2, 64]
    assert 'stem_conv.weight' in info['tensors'] and info['tensors']['stem_conv.weight'] == [64, 33, 3, 3]
    assert len(info['conv3x3']) == 1 + 2 * 4 and all(info['template']['rename'][k] == k for k in info['conv3x3'])
    assert not info['template']['unmatched_netspec_keys']
    assert 'def forward' in info['forward_source']
