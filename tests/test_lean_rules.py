"""CPU: the flat-board candidate-move filter (maru_b200/dropin/lean_rules.h, C ABI maru_rules_eval) against
the reference's Board::getEnableds(color, checkSeki=true) and Board::getTerritories(color) (reference
src/deepgo/native/cpp/Board.cpp:333-378, 872-1040, 1177-1631, compiled unchanged into oracle/_ref).
The two decide which moves Evaluator::evaluate hands to the search (Evaluator.cpp:60-68): every cell
must agree. Random playouts rarely build eyes, so besides them the sample contains dense late-game
boards, boards assembled from small enclosed shapes (eyes, nakade, seki-like pockets) and ko positions."""
import numpy as np
import pytest

import maru_b200


@pytest.fixture(scope='module')
def ref():
    import torch  # noqa: F401
    from oracle import ref as r
    if not r.available('board'):
        pytest.skip('oracle/_ref reference board library not present')
    return r


def snapshot(ref, b, w, h):
    col = np.array([[b.get_color(x, y) for x in range(w)] for y in range(h)], np.int8)
    ko = None
    for c in (ref.BLACK, ref.WHITE):
        kx, ky = b.get_ko(c)
        if kx >= 0:
            ko = (kx, ky, c)
    return col, ko


def check(ref, b, w, h, stats):
    col, ko = snapshot(ref, b, w, h)
    for color in (ref.BLACK, ref.WHITE):
        want_en = b.get_enableds(color, True).reshape(h, w).astype(np.uint8)
        want_te = b.get_territories(color).reshape(h, w).astype(np.int8)
        got_en, got_te = maru_b200.capi.rules_eval(col, color, ko)
        assert np.array_equal(got_en, want_en), ('enableds', w, h, color, np.argwhere(got_en != want_en)[:5].tolist(),
                                                 col.tolist())
        assert np.array_equal(got_te, want_te), ('territories', w, h, color, np.argwhere(got_te != want_te)[:5].tolist(),
                                                 col.tolist())
        plain = b.get_enableds(color, False).reshape(h, w).astype(np.uint8)
        stats['seki_vetoes'] += int((plain != want_en).sum())
        stats['territory_cells'] += int((want_te != 0).sum())
        stats['positions'] += 1


@pytest.mark.parametrize('w,h,n_pos', [(19, 19, 60), (13, 13, 80), (9, 9, 200), (9, 13, 60), (5, 7, 150), (7, 7, 200)])
def test_random_playouts_including_dense_endgames(built_lib, ref, w, h, n_pos):
    rng = np.random.default_rng(500 + 17 * w + h)
    stats = dict(seki_vetoes=0, territory_cells=0, positions=0)
    for i in range(n_pos):
        hi = 0.95 if i % 2 else 0.6                      # every second playout runs deep into the endgame
        moves = int(rng.integers(w * h // 8, int(w * h * hi)) * (2 if i % 4 == 3 else 1))
        b, _c = ref.random_playout(w, h, moves, rng, pass_prob=0.01)
        check(ref, b, w, h, stats)
    assert stats['positions'] == 2 * n_pos


def build_from_shapes(ref, w, h, rng):
    """A board tiled with walls of one colour enclosing small pockets, partly filled with opponent stones: the
    shapes the seki / nakade / decided-territory rules look at. Stones are placed through play(), so captures
    resolve exactly as in the reference."""
    b = ref.RefBoard(w, h)
    owner = np.zeros((h, w), np.int8)
    for y in range(h):
        for x in range(w):
            owner[y, x] = ref.BLACK if ((x // 4 + y // 4) % 2 == 0) else ref.WHITE
    cells = [(x, y) for y in range(h) for x in range(w)]
    rng.shuffle(cells)
    for (x, y) in cells:
        r = rng.random()
        if r < 0.62:
            b.play(x, y, int(owner[y, x]))
        elif r < 0.80:
            b.play(x, y, int(-owner[y, x]))
    return b


@pytest.mark.parametrize('w,h,n_pos', [(9, 9, 300), (13, 13, 120), (19, 19, 50), (6, 6, 300)])
def test_shape_rich_boards(built_lib, ref, w, h, n_pos):
    rng = np.random.default_rng(900 + w)
    stats = dict(seki_vetoes=0, territory_cells=0, positions=0)
    for _ in range(n_pos):
        check(ref, build_from_shapes(ref, w, h, rng), w, h, stats)
    # the sample must actually exercise the rules under test
    assert stats['seki_vetoes'] > 0 and stats['territory_cells'] > 0, stats


def test_ko_point_is_excluded_for_the_right_colour_only(built_lib, ref):
    b = ref.RefBoard(9, 9)
    a, o = ref.BLACK, ref.WHITE
    for (x, y) in [(1, 0), (0, 1), (1, 2)]:
        assert b.play(x, y, a) >= 0
    for (x, y) in [(2, 0), (3, 1), (2, 2), (1, 1)]:
        assert b.play(x, y, o) >= 0
    assert b.play(2, 1, a) == 1 and b.get_ko(o) == (1, 1)
    stats = dict(seki_vetoes=0, territory_cells=0, positions=0)
    check(ref, b, 9, 9, stats)
    col, ko = snapshot(ref, b, 9, 9)
    assert maru_b200.capi.rules_eval(col, o, ko)[0][1, 1] == 0 and maru_b200.capi.rules_eval(col, o, None)[0][1, 1] == 1
