"""The frame-synthesis oracle (oracle/synth.py), pinned where it restates third-party code: its cv2.warpPerspective restatement
against cv2 itself (bit-exact, u8, INTER_LINEAR, BORDER_CONSTANT), its Philox4x32-10 against the Random123 known-answer vectors,
and the product's host-side parameter derivation (deepcharuco_b200/synth.py: gpu_frame_params) against the oracle's."""
import numpy as np

from deepcharuco_b200 import synth
from oracle import synth as S


def _random_homography(rng, Wd, Hd):
    import cv2
    side = rng.uniform(0.3, 0.9) * 240
    ang = rng.uniform(0, 2 * np.pi)
    c, s = np.cos(ang), np.sin(ang)
    sq = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], np.float32) * (side / 2)
    dst = sq @ np.array([[c, -s], [s, c]], np.float32).T + np.array([Wd / 2 + rng.uniform(-.2, .2) * Wd, Hd / 2 + rng.uniform(-.2, .2) * Hd], np.float32)
    dst += rng.uniform(-0.08, 0.08, (4, 2)).astype(np.float32) * side
    src = np.array([[0, 0], [239, 0], [239, 239], [0, 239]], np.float32)
    return cv2.getPerspectiveTransform(src, dst.astype(np.float32))


def warp_cases():
    rng = np.random.default_rng(0)
    board = synth.board_render(240)
    noise = rng.integers(0, 256, (240, 240)).astype(np.uint8)
    cases = []
    for it in range(24):
        Wd, Hd = (320, 240) if it % 3 else (640, 480)
        M = _random_homography(rng, Wd, Hd)
        if it == 7:
            M = np.eye(3)                                               # every source coordinate an exact integer (weight-table corner case)
        if it == 9:
            M = np.array([[1, 0, 10.5], [0, 1, -3.25], [0, 0, 1.0]])
        if it == 11:
            Wd, Hd = 200, 72                                            # a width that is not a multiple of cv2's 64-column blocks
        cases.append((board if it % 2 == 0 else noise, M, (Wd, Hd)))
    return cases


def test_warp_restatement_is_bit_exact_with_cv2():
    import cv2
    for src, M, dsize in warp_cases():
        want = cv2.warpPerspective(src, M, dsize, flags=cv2.INTER_LINEAR)
        got = S.warp_perspective_u8(src, cv2.invert(np.asarray(M, np.float64))[1], dsize)
        assert np.array_equal(want, got)
    mask = S.warp_perspective_u8(src, cv2.invert(np.asarray(M, np.float64))[1], dsize, constant_src=255)
    assert np.array_equal(mask, cv2.warpPerspective(np.full_like(src, 255), M, dsize, flags=cv2.INTER_LINEAR))


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32-10."""
    h = lambda t: [int(v) for v in t]
    assert h(S.philox4x32(0, 0, 0, 0, 0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert h(S.philox4x32(f, f, f, f, f, f)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert h(S.philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    assert h(synth._philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_host_parameters_match_the_oracle():
    for (H, W, nb, seed) in ((240, 320, 1, 5), (480, 640, 4, 9)):
        p = synth.gpu_frame_params(6, H, W, seed, first_index=3)
        for i in range(6):
            q = S.frame_params(seed, 3 + i, H, W, nb)
            assert q["lat_step"] == p["lat_step"][i] and q["bg_lo"] == p["bg_lo"][i] and q["bg_hi"] == p["bg_hi"][i] and q["gain"] == p["gain"][i]
            assert np.array_equal(q["blur_w"], p["blur_w"][i]) and abs(float(q["blur_w"].sum()) - 1) < 1e-6
            assert np.allclose(np.array(q["Hinv"]), p["Hinv"][i], rtol=1e-9, atol=1e-10)
            assert np.allclose(np.array(q["corners"]), p["corners"][i], rtol=0, atol=1e-9)
    arr = synth.pack_frame_params(p, H, W)
    assert arr[2].lat_h == H // arr[2].lat_step + 3 and arr[2].n_boards == 4 and abs(arr[2].hinv[3][8] - p["Hinv"][2, 3, 2, 2]) == 0


def test_oracle_frames_carry_detectable_boards(states):
    """The recipe produces frames the reference algorithm finds corners in, close to the ground-truth corners."""
    import oracle
    board = synth.board_render(240)
    frames, corners = S.make_frames(board, 4, 240, 320, seed=5)
    ks, errs = [], []
    for f, c in zip(frames, corners):
        r = oracle.pipeline.infer_gray(states[0], states[1], f)
        ks.append(0 if r.size == 0 else r.shape[0])
        if r.size:
            errs.append(np.linalg.norm(r[:, :2] - c[0][r[:, 2].astype(int)], axis=1).mean())
    assert np.mean(ks) > 10 and np.mean(errs) < 1.0
    assert np.float32(1.0 / 255.0).view(np.uint32) == 0x3B808081


def test_resize_restatement_is_bit_exact_with_cv2_when_shrinking():
    import cv2
    from oracle.decode import resize_linear_u8
    rng = np.random.default_rng(0)
    for (Hs, Ws, C) in ((1920, 2560, 3), (480, 640, 3), (720, 1280, 1), (300, 421, 3), (241, 323, 1), (240, 320, 3)):
        src = rng.integers(0, 256, (Hs, Ws, C)).astype(np.uint8)
        if C == 1:
            src = src[..., 0]
        assert np.array_equal(resize_linear_u8(src, (320, 240)), cv2.resize(src, (320, 240), interpolation=cv2.INTER_LINEAR))
