"""Device pre-processing through the C ABI against the CPU path of the reference: expand2square
(mm_utils.py:128-139) + PIL bicubic resize + centre crop + /255 + normalise = oracle.clip_preprocess. Integer work -> bit-exact."""
import numpy as np
import pytest
import torch
from PIL import Image

from oracle import clipt5_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda", 0)


def rand_images(sizes, seed=0):
    rng = np.random.RandomState(seed)
    return [rng.randint(0, 256, (h, w, 3), dtype=np.uint8) for (w, h) in sizes]


SIZES = [(512, 512), (640, 400), (300, 500), (336, 336), (123, 77), (1600, 1200), (48, 900)]


@pytest.mark.parametrize("pad", [True, False])
def test_device_preprocess_bit_exact_vs_pil_path(dev, pad):
    from t2v_metrics_b200.engine import clip_preprocess_u8
    arrs = rand_images(SIZES, seed=1)
    got = clip_preprocess_u8([torch.from_numpy(a) for a in arrs], 336, dev, pad=pad)
    torch.cuda.synchronize()
    assert got.shape == (len(arrs), 3, 336, 336) and got.dtype == torch.float32
    for i, a in enumerate(arrs):
        ref = orc.clip_preprocess(Image.fromarray(a), 336, pad=pad)
        assert torch.equal(got[i].cpu(), ref), f"image {SIZES[i]} pad={pad}: max diff {float((got[i].cpu() - ref).abs().max())}"


def test_device_preprocess_smooth_image_and_bf16_and_device_input(dev):
    from t2v_metrics_b200.engine import clip_preprocess_u8
    # a smooth gradient + saturated blocks exercises the clip8 over/undershoot of the bicubic lobes
    y, x = np.mgrid[0:700, 0:900]
    a = np.stack([(x * 255 // 899), (y * 255 // 699), ((x // 50 + y // 50) % 2) * 255], axis=-1).astype(np.uint8)
    ref = orc.clip_preprocess(Image.fromarray(a), 336)
    t = torch.from_numpy(a)
    got = clip_preprocess_u8([t], 336, dev)
    assert torch.equal(got[0].cpu(), ref)
    got_dev = clip_preprocess_u8([t.to(dev)], 336, dev)                  # device-resident source
    assert torch.equal(got_dev[0].cpu(), ref)
    got16 = clip_preprocess_u8([t], 336, dev, out_dtype=torch.bfloat16)
    assert torch.equal(got16[0].cpu(), ref.bfloat16())
    # 224-pixel tower size (CLIP ViT-L/14) and a batch of equal-size images sharing one table
    arrs = rand_images([(512, 512)] * 5, seed=3)
    got224 = clip_preprocess_u8([torch.from_numpy(v) for v in arrs], 224, dev)
    for i, v in enumerate(arrs):
        assert torch.equal(got224[i].cpu(), orc.clip_preprocess(Image.fromarray(v), 224))


def test_scores_from_u8_images_match_scores_from_cpu_preprocessing(dev):
    """score_images_u8 (device pre-processing) == score_tensors on the CPU-pre-processed pixels, bit for bit."""
    import dataclasses
    from t2v_metrics_b200.config import ClipT5Config
    from t2v_metrics_b200.engine import ClipT5Engine
    ocfg = orc.ClipT5Config.tiny()
    sd = orc.make_synthetic_state_dict(ocfg, seed=0)
    cfg = ClipT5Config(**dataclasses.asdict(ocfg))
    eng = ClipT5Engine(cfg, dev)
    eng.load_state_dict(sd)
    inp = orc.make_synthetic_inputs(ocfg, 3, 12, seed=2)
    arrs = rand_images([(80, 60), (64, 64), (50, 90)], seed=4)
    cpu_pixels = torch.stack([orc.clip_preprocess(Image.fromarray(a), cfg.image_size) for a in arrs])
    ids, lens, labels = (inp[k].to(torch.int32) for k in ("input_ids", "text_lens", "labels"))
    a = eng.score_tensors(cpu_pixels.to(dev), ids.to(dev), lens.to(dev), labels.to(dev)).cpu()
    b = eng.score_images_u8([torch.from_numpy(v) for v in arrs], ids, lens, labels)
    assert torch.equal(a, b)


def test_qwen_device_preprocess_bit_exact(dev):
    """smart_resize + PIL bicubic + normalise + frame duplication + merge-order patch rows on the device == the CPU path
    (qwen_utils.qwen_image_to_patches <-> qwen_vl_utils + Qwen2VLImageProcessor, image_processing_qwen2_vl.py:62-87,191-220)."""
    from t2v_metrics_b200.engine import qwen_preprocess_u8
    from t2v_metrics_b200.models.vqascore_models.qwen_utils import qwen_image_to_patches
    sizes = [(448, 448), (640, 400), (300, 500), (123, 77), (30, 30), (1600, 1200), (2200, 1700)]
    arrs = rand_images(sizes, seed=5)
    got, grids = qwen_preprocess_u8([torch.from_numpy(a) for a in arrs], dev)
    torch.cuda.synchronize()
    row = 0
    for a, g in zip(arrs, grids):
        ref, gref = qwen_image_to_patches(Image.fromarray(a))
        assert tuple(g) == tuple(gref)
        n = ref.shape[0]
        assert torch.equal(got[row:row + n].cpu(), ref), f"grid {g}"
        row += n
    assert row == got.shape[0]
    got16, _ = qwen_preprocess_u8([torch.from_numpy(arrs[0])], dev, out_dtype=torch.bfloat16)
    assert torch.equal(got16.cpu(), qwen_image_to_patches(Image.fromarray(arrs[0]))[0].bfloat16())
