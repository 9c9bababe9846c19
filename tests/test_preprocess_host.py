"""Host logic of the device pre-processing, no GPU: the fixed-point bicubic tap tables libvqa_b200.so builds
(vqa_resample_table, the same code vqa_clip_preprocess uploads to the device) must reproduce Pillow's resize bit for bit when the
two integer passes are replayed in numpy, and the geometry must match the reference's CLIP processor (shortest edge -> S, centre crop)."""
import ctypes as C

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import clipt5_oracle as orc
from t2v_metrics_b200 import _lib

PREC = 22


def table(in_size, out_size, first=0, count=None):
    lib = _lib.load()
    count = out_size - first if count is None else count
    ks = lib.vqa_resample_table(in_size, out_size, first, count, None, None)
    assert ks > 0
    b = (C.c_int32 * (2 * count))()
    k = (C.c_int32 * (ks * count))()
    assert lib.vqa_resample_table(in_size, out_size, first, count, b, k) == ks
    return np.array(b, dtype=np.int64).reshape(count, 2), np.array(k, dtype=np.int64).reshape(count, ks)


def resample_axis(img, bounds, kk):
    """One pass along axis 1 of [rows, n, 3] uint8 with Resample.c's rounding: (sum + 2^21) >> 22 clipped to uint8."""
    out = np.empty((img.shape[0], bounds.shape[0], 3), dtype=np.uint8)
    src = img.astype(np.int64)
    for xx, (x0, n) in enumerate(bounds):
        acc = (src[:, x0:x0 + n, :] * kk[xx, :n][None, :, None]).sum(axis=1) + (1 << (PREC - 1))
        out[:, xx, :] = np.clip(acc >> PREC, 0, 255).astype(np.uint8)
    return out


def replay_resize(arr, nw, nh):
    h, w = arr.shape[:2]
    hb, hk = table(w, nw)
    vb, vk = table(h, nh)
    tmp = resample_axis(arr, hb, hk)                                        # horizontal first, rounded to uint8
    return resample_axis(tmp.transpose(1, 0, 2), vb, vk).transpose(1, 0, 2)  # then vertical


@pytest.mark.parametrize("w,h,nw,nh", [(512, 512, 336, 336), (640, 400, 537, 336), (300, 500, 336, 560), (200, 150, 448, 336),
                                       (336, 336, 336, 336), (1000, 37, 336, 12), (97, 1301, 25, 336)])
def test_tables_reproduce_pillow_resize(w, h, nw, nh):
    rng = np.random.RandomState(w * 7 + h)
    arr = rng.randint(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(arr).resize((nw, nh), resample=Image.BICUBIC))
    got = replay_resize(arr, nw, nh)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref), f"max diff {np.abs(got.astype(int) - ref.astype(int)).max()}"


def test_table_slices_and_identity():
    b, k = table(512, 336)
    b2, k2 = table(512, 336, 100, 50)
    assert np.array_equal(b[100:150], b2) and np.array_equal(k[100:150], k2)
    bi, ki = table(336, 336)
    for xx in range(336):                       # same size: a single tap of exactly 1.0 on the pixel itself
        x0, n = bi[xx]
        taps = {int(x0 + j): int(ki[xx, j]) for j in range(n) if ki[xx, j] != 0}
        assert taps == {xx: 1 << PREC}
    assert _lib.load().vqa_resample_table(0, 336, 0, 336, None, None) == 0
    assert "bad resample" in _lib.last_error(None)


@pytest.mark.parametrize("w,h", [(512, 512), (640, 400), (300, 500), (123, 77)])
def test_replayed_pipeline_equals_oracle_preprocess(w, h):
    """expand2square + resize + crop + normalise replayed with the library's tables == oracle.clip_preprocess (PIL path)."""
    rng = np.random.RandomState(w + h)
    arr = rng.randint(0, 256, (h, w, 3), dtype=np.uint8)
    ref = orc.clip_preprocess(Image.fromarray(arr), 336)
    side = max(w, h)
    canvas = np.empty((side, side, 3), dtype=np.uint8)
    canvas[:] = np.array([int(x * 255) for x in orc.CLIP_MEAN], dtype=np.uint8)
    canvas[(side - h) // 2:(side - h) // 2 + h, (side - w) // 2:(side - w) // 2 + w] = arr
    u8 = replay_resize(canvas, 336, 336)
    x = u8.astype(np.float32) / np.float32(255.0)
    x = (x - np.asarray(orc.CLIP_MEAN, dtype=np.float32)) / np.asarray(orc.CLIP_STD, dtype=np.float32)
    assert torch.equal(torch.from_numpy(x).permute(2, 0, 1), ref)


def test_workspace_sizing_and_errors():
    lib = _lib.load()
    H, W = (C.c_int32 * 3)(512, 400, 512), (C.c_int32 * 3)(512, 640, 512)
    n = lib.vqa_clip_preprocess_workspace_bytes(H, W, 3, 336, 1)
    assert n > 0
    one = lib.vqa_clip_preprocess_workspace_bytes((C.c_int32 * 1)(512), (C.c_int32 * 1)(512), 1, 336, 1)
    assert 0 < one < n                                              # tables are shared between images of the same canvas size
    assert lib.vqa_clip_preprocess_workspace_bytes((C.c_int32 * 1)(0), (C.c_int32 * 1)(5), 1, 336, 1) == 0
    # a 200000-pixel-high strip cannot keep its vertical window in shared memory: refused, not mangled
    assert lib.vqa_clip_preprocess_workspace_bytes((C.c_int32 * 1)(200000), (C.c_int32 * 1)(200000), 1, 336, 1) == 0
    assert "too large" in _lib.last_error(None)


def test_qwen_plan_matches_python_smart_resize():
    """vqa_qwen_preprocess_plan (C++ smart_resize incl. Python's round-half-even) == qwen_utils.smart_resize on a sweep of sizes,
    including the min-pixel and max-pixel branches and exact .5 ties (h / 28 = k + 0.5)."""
    from t2v_metrics_b200.engine import qwen_preprocess_plan
    from t2v_metrics_b200.models.vqascore_models.qwen_utils import smart_resize
    rng = np.random.RandomState(0)
    sizes = [(448, 448), (14, 14), (42, 70), (126, 98), (30, 30), (40, 2000), (3000, 4000), (1003, 1003), (2047, 777), (98, 154)]
    sizes += [(int(a), int(b)) for a, b in rng.randint(20, 2500, (60, 2))]
    grids, total, wsb = qwen_preprocess_plan(sizes)
    ref = [smart_resize(h, w) for h, w in sizes]
    assert [(g[1] * 14, g[2] * 14) for g in grids] == ref
    assert total == sum(h // 14 * (w // 14) for h, w in ref) and wsb > 0
    with pytest.raises(RuntimeError, match="aspect ratio"):
        qwen_preprocess_plan([(10, 2500)])
