"""Architecture descriptions for the engine (the HF config objects the reference loads with `from_pretrained`,
t2v_metrics/models/vqascore_models/mm_utils.py:182-241, reduced to what the kernels need)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class ClipT5Config:
    # CLIP ViT-L/14-336 vision tower
    image_size: int = 336
    patch_size: int = 14
    vit_hidden: int = 1024
    vit_heads: int = 16
    vit_mlp: int = 4096
    vit_layers: int = 24           # layers in the checkpoint; mm_vision_select_layer=-2 => vit_layers - 1 are executed
    vit_ln_eps: float = 1e-5
    # FlanT5 encoder-decoder
    d_model: int = 4096
    n_heads: int = 64
    d_kv: int = 64
    d_ff: int = 10240
    enc_layers: int = 24
    dec_layers: int = 24
    vocab: int = 32128
    rel_buckets: int = 32
    rel_max_distance: int = 128
    t5_ln_eps: float = 1e-6
    pad_token_id: int = 0
    decoder_start_id: int = 0

    @property
    def num_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @staticmethod
    def xxl() -> "ClipT5Config":
        """clip-flant5-xxl (google/flan-t5-xxl + openai/clip-vit-large-patch14-336), SURVEY App. A."""
        return ClipT5Config()

    @staticmethod
    def xl() -> "ClipT5Config":
        return ClipT5Config(d_model=2048, n_heads=32, d_ff=5120)


CLIPT5_MODELS = {
    # public names of the reference's (v3.0) model table; checkpoints: zhiqiulin/clip-flant5-{xxl,xl}
    "clip-flant5-xxl": dict(config=ClipT5Config.xxl, tokenizer="google/flan-t5-xxl", weights="zhiqiulin/clip-flant5-xxl"),
    "clip-flant5-xl": dict(config=ClipT5Config.xl, tokenizer="google/flan-t5-xl", weights="zhiqiulin/clip-flant5-xl"),
}


@dataclass
class Qwen25VLConfig:
    """Qwen2.5-VL dims (7B defaults; transformers configuration_qwen2_5_vl.py:51-64,106-125)."""
    vit_depth: int = 32
    vit_hidden: int = 1280
    vit_heads: int = 16
    vit_mlp: int = 3420
    patch_size: int = 14
    temporal_patch_size: int = 2
    spatial_merge_size: int = 2
    window_size: int = 112
    fullatt_block_indexes: tuple = (7, 15, 23, 31)
    out_hidden: int = 3584
    tokens_per_second: int = 2
    hidden: int = 3584
    layers: int = 28
    heads: int = 28
    kv_heads: int = 4
    mlp: int = 18944
    vocab: int = 152064
    rms_eps: float = 1e-6
    rope_theta: float = 1e6
    mrope_section: tuple = (16, 24, 24)
    image_token_id: int = 151655
    video_token_id: int = 151656

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def vit_head_dim(self) -> int:
        return self.vit_hidden // self.vit_heads

    @property
    def patch_dim(self) -> int:
        return 3 * self.temporal_patch_size * self.patch_size ** 2

    @property
    def vit_mlp_padded(self) -> int:
        return (self.vit_mlp + 127) // 128 * 128

    @staticmethod
    def qwen25_vl_7b() -> "Qwen25VLConfig":
        return Qwen25VLConfig()


QWEN25VL_MODELS = {
    "qwen2.5-vl-7b": dict(config=Qwen25VLConfig.qwen25_vl_7b, weights="Qwen/Qwen2.5-VL-7B-Instruct"),
}
