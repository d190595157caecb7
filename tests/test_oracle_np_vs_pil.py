"""Pin the NumPy arithmetic model (oracle/np_model.py) against Pillow itself, primitive by
primitive (the PROBE list of SURVEY.md 8c re-created in-repo)."""
import numpy as np
import PIL.Image
import PIL.ImageDraw
import PIL.ImageEnhance
import PIL.ImageFilter
import PIL.ImageOps
import pytest

from oracle import np_model as nm


def rnd(shape, seed):
    return np.random.default_rng(seed).integers(0, 256, shape + (3,), dtype=np.uint8)


def test_blend_exhaustive_pairs():
    """all 65,536 (a,b) byte pairs x 61 alphas in [0.1,1.9]: fp32 model == Image.blend"""
    a = np.repeat(np.arange(256, dtype=np.uint8), 256).reshape(256, 256)
    b = np.tile(np.arange(256, dtype=np.uint8), 256).reshape(256, 256)
    ia = np.stack([a, a, a], -1)
    ib = np.stack([b, b, b], -1)
    pa, pb = PIL.Image.fromarray(ia), PIL.Image.fromarray(ib)
    for alpha in np.linspace(0.1, 1.9, 61):
        want = np.asarray(PIL.Image.blend(pa, pb, float(alpha)))
        assert np.array_equal(nm.blend(ia, ib, float(alpha)), want), alpha


@pytest.mark.parametrize("shape", [(32, 32), (224, 224), (17, 63), (2, 9), (3, 3)])
def test_smooth_luma_contrast_sharpness(shape):
    for seed in range(6):
        img = rnd(shape, seed)
        p = PIL.Image.fromarray(img)
        assert np.array_equal(nm.smooth3x3(img), np.asarray(p.filter(PIL.ImageFilter.SMOOTH)))
        assert np.array_equal(nm.luma(img), np.asarray(p.convert("L")))
        for v in (0.1, 0.77, 1.0, 1.3, 1.9):
            assert np.array_equal(nm.contrast(img, v), np.asarray(PIL.ImageEnhance.Contrast(p).enhance(v)))
            assert np.array_equal(nm.sharpness(img, v), np.asarray(PIL.ImageEnhance.Sharpness(p).enhance(v)))
            assert np.array_equal(nm.color(img, v), np.asarray(PIL.ImageEnhance.Color(p).enhance(v)))
            assert np.array_equal(nm.brightness(img, v), np.asarray(PIL.ImageEnhance.Brightness(p).enhance(v)))


@pytest.mark.parametrize("shape", [(32, 32), (224, 224), (380, 380), (31, 45), (6, 4)])
def test_affine_models(shape):
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    img = rnd(shape, 3)
    p = PIL.Image.fromarray(img)
    for _ in range(25):
        v = float(rng.uniform(-0.3, 0.3))
        assert np.array_equal(nm.affine_nearest(img, (1, v, 0, 0, 1, 0)),
                              np.asarray(p.transform(p.size, PIL.Image.AFFINE, (1, v, 0, 0, 1, 0))))
        assert np.array_equal(nm.affine_nearest(img, (1, 0, 0, v, 1, 0)),
                              np.asarray(p.transform(p.size, PIL.Image.AFFINE, (1, 0, 0, v, 1, 0))))
        t = float(rng.uniform(-0.45, 0.45))
        assert np.array_equal(nm.affine_nearest(img, (1, 0, t * w, 0, 1, 0)),
                              np.asarray(p.transform(p.size, PIL.Image.AFFINE, (1, 0, t * w, 0, 1, 0))))
        assert np.array_equal(nm.affine_nearest(img, (1, 0, 0, 0, 1, t * h)),
                              np.asarray(p.transform(p.size, PIL.Image.AFFINE, (1, 0, 0, 0, 1, t * h))))
        ang = float(rng.uniform(-30, 30))
        assert np.array_equal(nm.rotate(img, ang), np.asarray(p.rotate(ang)))
    for k in (-10, -3.5, -0.5, 0, 0.49, 0.5, 1.0, 7.25, 10):
        assert np.array_equal(nm.affine_nearest(img, (1, 0, k, 0, 1, 0)),
                              np.asarray(p.transform(p.size, PIL.Image.AFFINE, (1, 0, k, 0, 1, 0))))
    assert np.array_equal(nm.rotate(img, 0.0), img)
    assert np.array_equal(nm.rotate(img, 360.0), img)


def test_histogram_luts_low_entropy_and_overflow():
    rng = np.random.default_rng(0)
    for i in range(120):
        s = int(rng.integers(4, 70))
        lo = int(rng.integers(0, 250))
        hi = int(rng.integers(lo, 256))
        img = rng.integers(lo, hi + 1, (s, s, 3), dtype=np.uint8)
        if i % 5 == 0:
            img[...] = img[0, 0]                   # constant colour
        if i % 7 == 0:                             # few distinct values, skewed: exercises LUT clipping
            img = (img // 64 * 64).astype(np.uint8)
            img[0, 0] = 255
        p = PIL.Image.fromarray(img)
        assert np.array_equal(nm.autocontrast(img), np.asarray(PIL.ImageOps.autocontrast(p)))
        assert np.array_equal(nm.equalize(img), np.asarray(PIL.ImageOps.equalize(p)))
    # crafted Equalize overflow: LUT entry > 255 must clip
    img = np.zeros((20, 20, 3), np.uint8)
    img[:, :10] = 10
    img[0, 0] = 200
    assert np.array_equal(nm.equalize(img), np.asarray(PIL.ImageOps.equalize(PIL.Image.fromarray(img))))


def test_autocontrast_every_lo_hi_pair():
    """all 32,640 (lo,hi) pairs: the fp64 mul-then-add LUT == PIL's"""
    for lo in range(0, 255, 1):
        his = np.arange(lo + 1, 256)
        # one image row per hi: pixels {lo, hi}
        img = np.zeros((len(his), 2, 3), np.uint8)
        img[:, 0] = lo
        img[:, 1] = his[:, None]
        if lo % 16:        # keep the runtime small: PIL call per row only for a subset
            continue
        for r, hi in enumerate(his):
            row = img[r:r + 1]
            assert np.array_equal(nm.autocontrast(row), np.asarray(PIL.ImageOps.autocontrast(PIL.Image.fromarray(row))))
    for lo in range(255):
        for hi in (lo + 1, min(255, lo + 37), 255):
            if hi <= lo:
                continue
            h = np.zeros(256, np.int64)
            h[lo] = 1
            h[hi] = 1
            lut = nm.autocontrast_lut(h)
            ramp = np.arange(256, dtype=np.uint8)[None, :, None].repeat(3, 2)
            ramp_img = np.concatenate([ramp, ramp], 0)
            ramp_img[1, 0] = lo
            # PIL reference: build an image whose histogram extremes are lo/hi and contains every value in between
            vals = np.arange(lo, hi + 1, dtype=np.uint8)
            im = np.stack([vals, vals, vals], -1)[None]
            want = np.asarray(PIL.ImageOps.autocontrast(PIL.Image.fromarray(im)))
            assert np.array_equal(np.asarray(lut, dtype=np.uint8)[vals], want[0, :, 0]), (lo, hi)


def test_point_ops():
    img = rnd((40, 40), 9)
    p = PIL.Image.fromarray(img)
    assert np.array_equal(nm.invert(img), np.asarray(PIL.ImageOps.invert(p)))
    for v in (0, 0.2, 17.0, 17.5, 128, 255.9, 256):
        assert np.array_equal(nm.solarize(img, v), np.asarray(PIL.ImageOps.solarize(p, v)))
    for v in (0, 0.9, 1, 3.99, 4, 5.6, 7.2, 8):
        assert np.array_equal(nm.posterize(img, v), np.asarray(PIL.ImageOps.posterize(p, int(v))))


def test_cutout_rectangle_inclusive_trunc():
    rng = np.random.default_rng(2)
    for _ in range(600):
        w, h = int(rng.integers(4, 60)), int(rng.integers(4, 60))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        v = float(rng.uniform(0, 0.2) * w) if rng.random() < 0.7 else float(rng.integers(0, 21))
        ux, uy = float(rng.random()), float(rng.random())
        cx, cy = w + (1.0 - w) * ux, h + (1.0 - h) * uy
        x0, y0 = int(max(0, cx - v / 2.0)), int(max(0, cy - v / 2.0))
        x1, y1 = min(w, x0 + v), min(h, y0 + v)
        p = PIL.Image.fromarray(img).copy()
        PIL.ImageDraw.Draw(p).rectangle((x0, y0, x1, y1), (125, 123, 114))
        assert np.array_equal(nm.cutout_abs(img, v, ux, uy), np.asarray(p))


def test_totensor_normalize_is_true_fp32_division():
    import torch
    from torchvision.transforms import transforms as T
    from oracle import pil_path
    img = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    for mean, std in ((pil_path.CIFAR_MEAN, pil_path.CIFAR_STD), (pil_path.IMAGENET_MEAN, pil_path.IMAGENET_STD)):
        want = T.Compose([T.ToTensor(), T.Normalize(mean, std)])(PIL.Image.fromarray(img)).numpy()
        assert np.array_equal(nm.to_tensor_normalize(img, mean, std), want)
