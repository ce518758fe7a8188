"""The reference's own known-answer tests, restated against the oracle (CPU only).

These pin the oracle for rows a1, a3, a8 (bit-exact) and a2 (weakly) of SURVEY.md section 8."""
import numpy as np

from oracle import prepost as P


def test_resize_longest_side_kat():
    # reference test/test_segmentation.cpp:15-46 (round-half-up of dim*scale in float32)
    assert P.resize_longest_side(13, 19, 26)[1:3] == (18, 26)
    assert P.resize_longest_side(13, 19, 10)[1:3] == (7, 10)
    assert P.resize_longest_side(19, 13, 26)[1:3] == (26, 18)
    assert P.resize_longest_side(19, 13, 10)[1:3] == (10, 7)
    # scale == 1 -> no resize, original view passes through (segmentation.cpp:63-70)
    assert P.resize_longest_side(1024, 1024, 1024)[0] is False
    assert P.resize_longest_side(1024, 700, 1024)[0] is False


def test_transform_kat():
    # reference test/test_segmentation.cpp:48-57
    _, w, h, scale = P.resize_longest_side(10, 10, 20)
    assert (w, h) == (20, 20)
    assert (P.transform_coord(0, scale), P.transform_coord(0, scale)) == (0, 0)
    assert (P.transform_coord(10, scale), P.transform_coord(10, scale)) == (20, 20)
    assert (P.transform_coord(2, scale), P.transform_coord(7, scale)) == (4, 14)


def test_create_image_tensor_kat():
    # reference test/test_segmentation.cpp:59-83: 8x6 iota image, four channel orders
    expected = {P.RGB: [0, 1, 2, 3, 4, 24], P.RGBA: [0, 1, 2, 4, 5, 32], P.BGRA: [2, 1, 0, 6, 5, 34],
                P.ARGB: [1, 2, 3, 5, 6, 33]}
    for ch, exp in expected.items():
        bpp = P.bytes_per_pixel(ch)
        img = (np.arange(8 * 6 * bpp) % 256).astype(np.uint8).reshape(6, 8, bpp)
        t = P.create_image_tensor(img, ch)
        assert t.shape == (6, 8, 3) and t.dtype == np.float32
        got = [t[0, 0, 0], t[0, 0, 1], t[0, 0, 2], t[0, 1, 0], t[0, 1, 1], t[1, 0, 0]]
        assert got == [float(v) for v in exp]
    m = (np.arange(48) % 256).astype(np.uint8).reshape(6, 8, 1)
    t = P.create_image_tensor(m, P.MASK)
    assert (t[..., 0] == t[..., 1]).all() and (t[..., 1] == t[..., 2]).all() and t[1, 0, 0] == 8.0


def test_write_mask_image_kat():
    # reference test/test_segmentation.cpp:85-99: strictly > 0, row stride from the tensor (5), extent 4x2
    vals = np.array([0.0, 0.0, 0.2, -3.1, 0.0, 5.5, 0.0, 0.7, 0.0, 0.9], np.float32).reshape(1, 1, 2, 5)
    m = P.write_mask_image(vals, 0, 4, 2)
    assert m.tolist() == [[0, 0, 255, 0], [255, 0, 255, 0]]


def test_image_resize_kat():
    # reference test/test_image.cpp:51-69 (weak pin of the stb restatement: half-pixel centres + clamped edges)
    img = np.zeros((8, 8, 4), np.uint8)
    for i in range(64):
        img[i // 8, i % 8] = [255, 4 * (i // 8), 4 * (i % 8), 255]
    r = P.resize_srgb(img, 4, 4)
    for i in range(16):
        assert r[i // 4, i % 4].tolist() == [255, 2 + 8 * (i // 4), 2 + 8 * (i % 4), 255]


def test_resize_is_identity_preserving_and_strided():
    rng = np.random.default_rng(0)
    flat = np.full((20, 20, 3), 77, np.uint8)
    assert (P.resize_srgb(flat, 9, 7) == 77).all()  # weights sum to one
    assert (P.resize_srgb(flat, 41, 33) == 77).all()
    # strided input is honoured (image.cpp:38-42)
    img = rng.integers(0, 256, (16, 12, 3), dtype=np.uint8)
    buf = np.zeros((16, 12 * 3 + 10), np.uint8)
    buf[:, :36] = img.reshape(16, 36)
    view = np.lib.stride_tricks.as_strided(buf, (16, 12, 3), (46, 3, 1))
    assert (P.resize_srgb(view, 7, 9, stride=46) == P.resize_srgb(img, 7, 9)).all()


def test_prompt_tensors():
    # reference segmentation.cpp:134-152 on the truck fixture: 1800x1200 -> scale 1024/1800; (486,722)->(276,411)
    _, w, h, scale = P.resize_longest_side(1800, 1200)
    assert (w, h) == (1024, 683)
    c, l = P.prompt_tensors((486, 722), None, scale)
    assert c.tolist() == [[[276.0, 411.0], [0.0, 0.0]]] and l.tolist() == [[1.0, -1.0]]
    c, l = P.prompt_tensors(None, (180, 110, 505, 330), np.float32(2.0))
    assert c.tolist() == [[[360.0, 220.0], [1010.0, 660.0]]] and l.tolist() == [[2.0, 3.0]]
