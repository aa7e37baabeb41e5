"""GPU: the BASELINE.json sizes (b4c128 @ batch 256, b10c256 @ batch 512) through size-independent
properties — the oracle only finishes small batches in seconds, so at full size the checks are:
determinism, batch-permutation equivariance (a position's result does not depend on its neighbours
in the batch or on its slot), ragged sub-batches equal the prefix of the full batch, probability
planes are normalised on the board and exactly zero off it, and the compact transport agrees."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CELLS = 361


def mixed_states(n: int, seed: int) -> np.ndarray:
    from maru_b200 import synth
    parts = [synth.random_states(n - 2 * (n // 4), 19, seed=seed),
             synth.random_states(n // 4, 13, seed=seed + 1),
             synth.random_states(n // 4, 9, seed=seed + 2)]
    st = np.concatenate(parts)
    return st[np.random.default_rng(seed).permutation(len(st))]


@pytest.mark.parametrize('name,batch', [('b4c128', 256), ('b10c256', 512)])
def test_full_size_properties(built_lib, model_dir, name, batch):
    import maru_b200
    ev = maru_b200.Evaluator(model_dir(name), gpu=0, max_batch=batch)
    try:
        st = mixed_states(batch, seed=11)
        out = ev.forward_states(st)
        launches = ev.info().launches_per_forward
        assert out.shape == (batch, 2169) and np.isfinite(out).all()
        assert np.array_equal(ev.forward_states(st), out)                     # run-to-run deterministic
        perm = np.random.default_rng(3).permutation(batch)
        assert np.array_equal(ev.forward_states(st[perm]), out[perm])         # slot / neighbour independent
        for n in (1, 7, batch - 1):
            got = ev.forward_states(st[:n])                                   # ragged prefix of the same bucket
            if ev.info().launches_per_forward == launches:
                assert np.array_equal(got, out[:n])
            else:
                # The engine picks the persistent multi-layer conv kernels or one launch per layer by the LIVE batch
                # size (Engine::flow_pays). Same MMAs in the same order, but the residual is added in the epilogue
                # (fp32) instead of by an identity MMA: one bf16 rounding may differ per residual layer.
                assert np.abs(got - out[:n]).max() < 2e-3
        # on-board mask from the records: width x height centred in the 19x19 frame
        w, h = st[:, 368].astype(int), st[:, 369].astype(int)
        yy, xx = np.mgrid[0:19, 0:19]
        ox, oy = (19 - w) // 2, (19 - h) // 2
        on = ((xx[None] >= ox[:, None, None]) & (xx[None] < (ox + w)[:, None, None]) &
              (yy[None] >= oy[:, None, None]) & (yy[None] < (oy + h)[:, None, None])).reshape(batch, CELLS)
        planes = out[:, :6 * CELLS].reshape(batch, 6, CELLS)
        assert (planes[:, :, :][~np.broadcast_to(on[:, None, :], planes.shape)] == 0).all()   # exactly 0 off board
        assert np.abs(planes[:, 0].sum(1) - 1).max() < 1e-4 and np.abs(planes[:, 1].sum(1) - 1).max() < 1e-4
        terr = planes[:, 2:5].sum(1)
        assert np.abs(terr[on] - 1).max() < 1e-5
        assert ((planes >= 0) & (planes <= 1)).all()
        assert ((out[:, 2166] > 0) & (out[:, 2166] < 1)).all()
        # compact transport: same bits as plane 0 in board order + the scalars
        leaf = ev.forward_leaves(st)
        for i in (0, batch // 2, batch - 1):
            sel = planes[i, 0][on[i]]                                        # raster order of the board cells
            assert np.array_equal(leaf[i, :w[i] * h[i]], sel)
            assert np.array_equal(leaf[i, 361:], out[i, 2166:])
    finally:
        ev.close()
