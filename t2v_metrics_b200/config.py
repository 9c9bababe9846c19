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
