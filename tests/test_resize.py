"""Image resize in front of the transforms.  Three layers, all bit-exact:
oracle/resize.py (numpy restatement of Pillow's 8-bit bicubic resampler) == Pillow itself (what the reference calls through
torchvision F.resize, data/transforms.py:87); the C library's host-side weight tables == the oracle's; and on the GPU the
device result == Pillow."""
import ctypes as C

import numpy as np
import pytest
from PIL import Image

from oracle import resize as oresize

CASES = [((70, 98), (64, 96)), ((448, 448), (448, 448)), ((600, 800), (448, 588)), ((100, 130), (224, 294)),
         ((900, 1200), (768, 1024)), ((33, 47), (28, 42)), ((512, 512), (980, 980)), ((1400, 300), (980, 210)),
         ((17, 17), (32, 32)), ((300, 300), (14, 14)), ((448, 600), (448, 448)), ((600, 448), (448, 448))]


def _img(shape, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
    a[: shape[0] // 3] = 255 * (np.indices(a[: shape[0] // 3].shape)[1] % 2)         # hard edges: over/undershoot -> clipping
    return a


@pytest.mark.parametrize("src,dst", CASES)
def test_oracle_equals_pillow(src, dst):
    a = _img(src, src[0] + dst[1])
    ref = np.asarray(Image.fromarray(a).resize((dst[1], dst[0]), Image.BICUBIC))
    assert np.array_equal(oresize.resize_bicubic_u8(a, *dst), ref)


@pytest.mark.parametrize("n_in,n_out", [(98, 96), (800, 588), (130, 294), (1200, 1024), (300, 14), (17, 32), (2000, 980), (5, 980)])
def test_library_weight_tables_equal_oracle(n_in, n_out):
    from unimedvl_b200 import _lib
    lib = _lib.load()
    ks = C.c_int32()
    _lib.check(lib.umv_resize_coefficients(n_in, n_out, None, None, C.byref(ks)))
    bounds = np.zeros((n_out, 2), dtype=np.int32)
    kk = np.zeros((n_out, ks.value), dtype=np.int32)
    _lib.check(lib.umv_resize_coefficients(n_in, n_out, bounds.ctypes.data_as(C.POINTER(C.c_int32)),
                                           kk.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(ks)))
    ob, ok = oresize.coefficients(n_in, n_out)
    assert ks.value == ok.shape[1] and np.array_equal(bounds, ob) and np.array_equal(kk, ok)


@pytest.fixture(scope="module")
def eng():
    from unimedvl_b200 import config as ucfg
    from unimedvl_b200.engine import Engine
    return Engine(ucfg.tiny(), max_tokens=64, max_seqs=2, kv_pages=8)


@pytest.mark.gpu
@pytest.mark.parametrize("src,dst", CASES)
def test_device_resize_equals_pillow(eng, src, dst):
    import torch
    a = _img(src, src[0] + dst[1])
    ref = np.asarray(Image.fromarray(a).resize((dst[1], dst[0]), Image.BICUBIC))
    got = eng.resize_u8(torch.from_numpy(a), *dst).cpu().numpy()
    assert np.array_equal(got, ref)


@pytest.mark.gpu
def test_device_resize_then_patchify_equals_host_transform(eng):
    """The whole ViT pre-processing chain on the device (resize rule -> resize -> normalise -> patchify) against the reference's
    host chain restated in packing.ImageTransform + patchify (bit-exact vs the reference fixtures in tests/test_packing.py)."""
    import torch
    from unimedvl_b200 import packing
    dims = eng.dims
    tf = packing.ImageTransform(980, 28, 14)
    for k, shape in enumerate([(100, 150), (333, 211), (40, 40)]):
        a = _img(shape, 90 + k)
        t = tf(Image.fromarray(a))
        want = packing.patchify(t, dims.vit.patch)
        w, h = tf.resize_transform.target_size(shape[1], shape[0])
        sized = eng.resize_u8(torch.from_numpy(a), h, w)
        pixels, pos, lens = eng.patchify_u8([sized])
        assert lens == [want.shape[0]] and torch.equal(pixels.cpu(), want)
        assert torch.equal(pos.cpu(), packing.flattened_position_ids(t.size(1), t.size(2), dims.vit.patch, dims.vit_max_num_patch_per_side))


def test_random_geometries_property():
    """Property over random geometries (seeded): oracle == Pillow and library weight tables == oracle, including extreme
    aspect ratios, 1-pixel sides and up/down scaling mixed per axis."""
    from unimedvl_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(2024)
    for _ in range(40):
        H, W, h, w = (int(v) for v in rng.integers(1, 160, 4))
        a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(a).resize((w, h), Image.BICUBIC))
        assert np.array_equal(oresize.resize_bicubic_u8(a, h, w), ref), (H, W, h, w)
        for n_in, n_out in ((W, w), (H, h)):
            ks = C.c_int32()
            _lib.check(lib.umv_resize_coefficients(n_in, n_out, None, None, C.byref(ks)))
            bounds = np.zeros((n_out, 2), dtype=np.int32)
            kk = np.zeros((n_out, ks.value), dtype=np.int32)
            _lib.check(lib.umv_resize_coefficients(n_in, n_out, bounds.ctypes.data_as(C.POINTER(C.c_int32)),
                                                   kk.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(ks)))
            ob, ok = oresize.coefficients(n_in, n_out)
            assert np.array_equal(bounds, ob) and np.array_equal(kk, ok), (n_in, n_out)
