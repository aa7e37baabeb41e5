"""CPU: whole games on the flat host board (maru_b200/dropin/lean_game.h, C ABI maru_game_*) against the
reference Board (reference src/deepgo/native/cpp/Board.cpp, compiled unchanged into oracle/_ref): the same
random move sequence — stones, captures, illegal tries, passes, kos — is played on both, and at every ply the
return value of play() and all 448 bytes of the compact record (cells with liberties and ladder flags,
histories, ko, move mask with the seki veto and decided territory) must be identical."""
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


@pytest.mark.parametrize('w,h,n_games,plies', [(9, 9, 25, 110), (19, 19, 6, 330), (13, 13, 10, 200), (5, 7, 40, 80),
                                               (7, 5, 20, 80)])
def test_games_move_by_move(built_lib, ref, w, h, n_games, plies):
    rng = np.random.default_rng(31 * w + h)
    n_illegal = n_pass = n_capture = n_ko = 0
    for g in range(n_games):
        rb = ref.RefBoard(w, h)
        lg = maru_b200.capi.LeanGame(w, h)
        color = ref.BLACK
        for ply in range(plies):
            r = rng.random()
            if r < 0.04:
                x, y = -1, -1                                   # pass
            else:
                x, y = int(rng.integers(0, w)), int(rng.integers(0, h))      # may be illegal or occupied
            a, b = rb.play(x, y, color), lg.play(x, y, color)
            assert a == b, (g, ply, x, y, color, a, b)
            n_illegal += a < 0
            n_pass += x < 0
            n_capture += a > 0
            if a >= 0:
                color = -color
            if ply % 3 == 0 or a > 0:                           # the record for the side to move, mask included
                komi, rule, sko = (7.5, ref.RULE_CH, False) if ply % 2 else (6.5, ref.RULE_JP, True)
                want = rb.get_state(color, komi, rule, sko, True)
                got = lg.state(color, komi, rule, sko, True)
                assert np.array_equal(got, want), (g, ply, np.flatnonzero(got != want)[:8].tolist())
                n_ko += int(want[373] != 0xFF)
        lg.close()
    assert n_illegal > 0 and n_pass > 0 and n_capture > 0
    if (w, h) in ((9, 9), (5, 7)):
        assert n_ko > 0                                         # ko points seen by the side to move


def test_clone_is_independent(built_lib):
    g = maru_b200.capi.LeanGame(9, 9)
    assert g.play(2, 2, 1) == 0 and g.play(2, 2, -1) == -1
    c = g.copy()
    assert c.play(3, 3, -1) == 0
    assert not np.array_equal(c.state(1), g.state(1))
    assert np.array_equal(g.copy().state(1), g.state(1))
    with pytest.raises(maru_b200.MaruB200Error):
        maru_b200.capi.LeanGame(20, 9)
