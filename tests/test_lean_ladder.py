"""CPU: the flat-board ladder read-out (maru_b200/dropin/lean_ladder.h, C ABI maru_ladder_flags) against
the reference's own Board::isShicho (Board::_updateShicho / _isShichoRen, reference
src/deepgo/native/cpp/Board.cpp:1045-1153, compiled unchanged into oracle/_ref). Ladder flags feed
feature planes 2 and 15, so the two must agree on every stone of every position."""
import numpy as np
import pytest

import maru_b200


@pytest.fixture(scope='module')
def ref():
    import torch  # noqa: F401  (the reference library needs torch loaded first)
    from oracle import ref as r
    if not r.available('board'):
        pytest.skip('oracle/_ref reference board library not present')
    return r


def colors_and_ko(ref, b, w, h):
    col = np.array([[b.get_color(x, y) for x in range(w)] for y in range(h)], np.int8)
    ko = None
    for c in (ref.BLACK, ref.WHITE):
        kx, ky = b.get_ko(c)
        if kx >= 0:
            ko = (kx, ky, c)
    return col, ko


def reference_flags(b, col, w, h):
    return np.array([[1 if (col[y, x] != 0 and b.is_shicho(x, y)) else 0 for x in range(w)] for y in range(h)],
                    np.uint8)


@pytest.mark.parametrize('w,h,n_pos', [(19, 19, 150), (13, 13, 150), (9, 9, 250), (9, 13, 80), (5, 7, 80)])
def test_flags_equal_reference_on_random_playouts(built_lib, ref, w, h, n_pos):
    rng = np.random.default_rng(1000 + 31 * w + h)
    n_atari = n_ladder = 0
    for _ in range(n_pos):
        moves = int(rng.integers(w * h // 6, int(w * h * 0.75)))
        b, _c = ref.random_playout(w, h, moves, rng)
        col, ko = colors_and_ko(ref, b, w, h)
        want = reference_flags(b, col, w, h)
        got = maru_b200.capi.ladder_flags(col, ko)
        assert np.array_equal(got, want), (w, h, moves, np.argwhere(got != want)[:5].tolist())
        n_ladder += int(want.sum() > 0)
        n_atari += int(any(col[y, x] != 0 and b.get_ren_space(x, y) == 1 for y in range(h) for x in range(w)))
    assert n_atari >= n_pos // 4 and n_ladder >= 5, (n_atari, n_ladder)    # the sample exercises the read-out


def test_ko_state_matters_and_matches(built_lib, ref):
    """Positions right after a ko capture: the read-out's first moves can be ko-illegal (Board.cpp:1121)."""
    n = 0
    rng = np.random.default_rng(9)
    for (w, h) in ((9, 9), (13, 13), (19, 19), (9, 13)):
        for (dx, dy) in [(0, 0), (w - 4, 0), (0, h - 3), (w - 4, h - 3), ((w - 4) // 2, (h - 3) // 2)]:
            for first in (ref.BLACK, ref.WHITE):
                # some random stones elsewhere first, then the ko shape of oracle/ref.py:ko_positions
                b = ref.RefBoard(w, h)
                a, o = first, -first
                for _ in range(int(rng.integers(0, w * h // 3))):
                    x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
                    if not (dx - 1 <= x <= dx + 4 and dy - 1 <= y <= dy + 3):
                        b.play(x, y, a if rng.random() < 0.5 else o)
                ok = all(b.play(dx + x, dy + y, a) >= 0 for (x, y) in [(1, 0), (0, 1), (1, 2)])
                ok = ok and all(b.play(dx + x, dy + y, o) >= 0 for (x, y) in [(2, 0), (3, 1), (2, 2), (1, 1)])
                if not ok or b.play(dx + 2, dy + 1, a) != 1 or b.get_ko(o) != (dx + 1, dy + 1):
                    continue
                col, ko = colors_and_ko(ref, b, w, h)
                assert ko == (dx + 1, dy + 1, o)
                want = reference_flags(b, col, w, h)
                assert np.array_equal(maru_b200.capi.ladder_flags(col, ko), want)
                n += 1
    assert n >= 20


def test_textbook_ladder(built_lib):
    # a white stone in atari that runs into the edge of a 9x9 board is a working ladder for Black; with a
    # white ladder breaker on the path it is not
    col = np.zeros((9, 9), np.int8)
    col[3, 3] = -1
    for x, y in ((3, 2), (2, 3), (3, 4), (4, 2)):   # black stones: the only liberty is (4,3), the ladder runs down-right
        col[y, x] = 1
    got = maru_b200.capi.ladder_flags(col)
    assert got[3, 3] == 1 and got.sum() == 1
    col[5, 6] = -1                                  # white ladder breaker on the diagonal
    assert maru_b200.capi.ladder_flags(col).sum() == 0
    with pytest.raises(maru_b200.MaruB200Error):
        maru_b200.capi.ladder_flags(np.zeros((20, 20), np.int8))


def test_compact_record_with_lean_ladder_is_byte_identical(built_lib, ref):
    """state_from_board(..., lean_ladder=true) — what the substitute Evaluator.cpp ships to the GPU — equals
    the record built with Board::isShicho, byte for byte, on fresh (ladder cache cold) copies."""
    rng = np.random.default_rng(77)
    for (w, h, n_pos) in ((19, 19, 40), (9, 9, 60), (13, 9, 30)):
        for _ in range(n_pos):
            b, c = ref.random_playout(w, h, int(rng.integers(5, int(w * h * 0.7))), rng)
            want = b.copy().get_state(c, 7.5, ref.RULE_CH, False, True)
            got = b.copy().get_state(c, 7.5, ref.RULE_CH, False, True, lean_ladder=True)
            assert np.array_equal(got, want)
