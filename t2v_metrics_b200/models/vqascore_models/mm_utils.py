"""Host-side pieces of the CLIP-FlanT5 path that survive in the reference's mm_utils.py, rebuilt for this engine."""
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
from PIL import Image

from ...constants import IMAGE_TOKEN_INDEX, DEFAULT_IMAGE_TOKEN

CLIP_IMAGE_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_IMAGE_STD = (0.26862954, 0.26130258, 0.27577711)


def expand2square(pil_img: Image.Image, background_color) -> Image.Image:
    """Centre the image on a square canvas of `background_color` (reference mm_utils.py:128-139)."""
    w, h = pil_img.size
    if w == h:
        return pil_img
    side = max(w, h)
    canvas = Image.new(pil_img.mode, (side, side), background_color)
    canvas.paste(pil_img, ((side - w) // 2, (side - h) // 2))
    return canvas


def t5_tokenizer_image_token(prompt: str, tokenizer, image_token_index: int = IMAGE_TOKEN_INDEX, return_tensors=None,
                             chunk_cache: Optional[dict] = None):
    """Tokenise every text chunk around '<image>' separately (each keeps its own trailing </s>) and join the chunks
    with `image_token_index` (reference mm_utils.py:164-179).
    `chunk_cache`: an exact memo chunk-string -> ids. The chunk before '<image>' is the constant system prompt and
    M x N scoring repeats every caption M times, so the slow SentencePiece tokenizer (`use_fast=False`, mm_utils.py:198) runs once
    per distinct chunk; whole chunks are the cache unit because sub-word tokenisation is not prefix-stable."""
    input_ids: List[int] = []
    for i, chunk in enumerate(prompt.split(DEFAULT_IMAGE_TOKEN)):
        if i > 0:
            input_ids.append(image_token_index)
        if chunk_cache is None:
            input_ids.extend(tokenizer(chunk).input_ids)
        else:
            ids = chunk_cache.get(chunk)
            if ids is None:
                ids = chunk_cache[chunk] = tuple(tokenizer(chunk).input_ids)
            input_ids.extend(ids)
    if return_tensors is not None:
        if return_tensors == 'pt':
            return torch.tensor(input_ids, dtype=torch.long)
        raise ValueError(f'Unsupported tensor type: {return_tensors}')
    return input_ids


def clip_preprocess(pil_img: Image.Image, image_size: int, pad: bool = True) -> torch.Tensor:
    """'pad' aspect mode of the v3.0 wrapper + CLIPImageProcessor (transformers image_processing_clip.py:23-34):
    expand2square(mean colour) -> bicubic resize (shortest edge) -> centre crop -> /255 -> normalise. fp32 [3,S,S]."""
    img = pil_img.convert("RGB")
    if pad:
        img = expand2square(img, tuple(int(x * 255) for x in CLIP_IMAGE_MEAN))
    w, h = img.size
    if w <= h:
        nw, nh = image_size, int(h * image_size / w)
    else:
        nw, nh = int(w * image_size / h), image_size
    if (nw, nh) != (w, h):
        img = img.resize((nw, nh), resample=Image.BICUBIC)
    w, h = img.size
    left, top = (w - image_size) // 2, (h - image_size) // 2
    img = img.crop((left, top, left + image_size, top + image_size))
    arr = np.asarray(img, dtype=np.float32) / 255.0
    arr = (arr - np.asarray(CLIP_IMAGE_MEAN, dtype=np.float32)) / np.asarray(CLIP_IMAGE_STD, dtype=np.float32)
    return torch.from_numpy(arr).permute(2, 0, 1).contiguous()
