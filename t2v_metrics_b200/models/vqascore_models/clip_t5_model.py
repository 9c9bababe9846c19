"""CLIP-FlanT5 VQAScore plugin backed by the B200 engine.

Same plugin contract as the reference's (v3.0) `CLIPT5Model(VQAScoreModel)` -- class attributes `video_mode`,
`allows_image`, `forward(images, texts, question_template, answer_template) -> CPU fp32 Tensor[n]` -- which is what
`Score.forward` calls once per image (reference t2v_metrics/score.py:104-106). The forward itself is one call into
libvqa_b200.so (engine.ClipT5Engine.score_tensors); nothing on this path runs through transformers or torch ops.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ...config import CLIPT5_MODELS as _MODEL_TABLE, ClipT5Config
from ...constants import HF_CACHE_DIR, CONTEXT_LEN, SYSTEM_MSG, IGNORE_INDEX, DEFAULT_IMAGE_TOKEN, IMAGE_TOKEN_INDEX
from .mm_utils import t5_tokenizer_image_token, clip_preprocess
from .vqa_model import VQAScoreModel

default_question_template = 'Does this figure show "{}"? Please answer yes or no.'  # V_3.0_README.md:211-215
default_answer_template = "Yes"

CLIP_T5_MODELS: Dict[str, dict] = {
    name: dict(tokenizer=dict(path=spec["tokenizer"], model_max_length=CONTEXT_LEN, padding_side="right"),
               model=dict(path=spec["weights"], conversation="t5_chat", image_aspect_ratio="pad"),
               config=spec["config"])
    for name, spec in _MODEL_TABLE.items()
}


def format_question(question: str, conversation_style: str = "t5_chat") -> str:
    """v3.0 format_question: SYSTEM_MSG + ' USER: <image>\\n' + question + ' ASSISTANT: ' (SURVEY App. A)."""
    if conversation_style == "t5_chat":
        return SYSTEM_MSG + " USER: " + DEFAULT_IMAGE_TOKEN + "\n" + question + " ASSISTANT: "
    if conversation_style == "t5_plain":
        return DEFAULT_IMAGE_TOKEN + "\n" + question
    raise NotImplementedError(conversation_style)


def format_answer(answer: str, conversation_style: str = "t5_chat") -> str:
    return answer


class CLIPT5Model(VQAScoreModel):
    """VQAScore with CLIP-FlanT5 on the B200 engine."""
    video_mode = "concat"
    allows_image = True

    def __init__(self, model_name="clip-flant5-xxl", device="cuda", cache_dir=HF_CACHE_DIR, tokenizer=None,
                 state_dict: Optional[Dict[str, torch.Tensor]] = None, config: Optional[ClipT5Config] = None,
                 checkpoint: Optional[str] = None, vision_tower_checkpoint: Optional[str] = None, cuda_graph_max_pairs: int = 16, **kwargs):
        assert model_name in CLIP_T5_MODELS
        # calls with at most this many pairs replay a CUDA graph captured per (pairs, images, padded text length, answer length): the
        # ~700 launches of a forward cost more host time than device time at batch 1 (0 disables)
        self.cuda_graph_max_pairs = int(cuda_graph_max_pairs)
        self._vision_checkpoint = vision_tower_checkpoint
        self._tokenizer_override = tokenizer
        self._state_dict = state_dict
        self._config_override = config
        self._checkpoint = checkpoint
        super().__init__(model_name=model_name, device=device, cache_dir=cache_dir)

    # ---- loading -------------------------------------------------------------------------------------------------
    def load_model(self):
        """Tokenizer via transformers (slow SentencePiece tokenizer, as mm_utils.py:198), weights from a local
        HF-named state dict (`state_dict=` / `checkpoint=` safetensors or .pt) cast to bf16 on the device
        (mm_utils.py:228) and bound into the engine."""
        from ...engine import ClipT5Engine

        spec = CLIP_T5_MODELS[self.model_name]
        self.conversational_style = spec["model"]["conversation"]
        self.image_aspect_ratio = spec["model"]["image_aspect_ratio"]
        self.context_len = spec["tokenizer"]["model_max_length"]
        self.cfg: ClipT5Config = self._config_override or spec["config"]()
        if self._tokenizer_override is not None:
            self.tokenizer = self._tokenizer_override
        else:
            from transformers import AutoTokenizer
            self.tokenizer = AutoTokenizer.from_pretrained(spec["tokenizer"]["path"], use_fast=False, cache_dir=self.cache_dir,
                                                           model_max_length=self.context_len, padding_side="right")
        sd = self._state_dict
        if sd is None:
            sd = self._load_checkpoint(self._checkpoint or spec["model"]["path"])
        dev = torch.device(self.device if self.device != "cuda" else "cuda:0")
        self.engine = ClipT5Engine(self.cfg, dev)
        self.engine.load_state_dict(sd)
        self._state_dict = None

    def _load_checkpoint(self, path: str) -> Dict[str, torch.Tensor]:
        """`path`: a local weights file or a downloaded repository directory (index json + shards) of zhiqiulin/clip-flant5-*;
        `vision_tower_checkpoint=`: the CLIP tower the reference loads separately when the checkpoint does not carry it
        (openai/clip-vit-large-patch14-336, mm_utils.py:226-227)."""
        from ...checkpoint import load_state_dict, normalise_clipt5_keys
        vision = load_state_dict(self._vision_checkpoint) if self._vision_checkpoint else None
        return normalise_clipt5_keys(load_state_dict(path), vision)

    # ---- pre-processing ------------------------------------------------------------------------------------------
    def load_images(self, image: List[str]) -> torch.Tensor:
        """image_loader (PIL decode, host) -> uint8 HWC -> ONE device kernel doing expand2square(mean colour) + PIL-exact bicubic
        resize + centre crop + /255 + normalise (engine.clip_preprocess_u8; bit-identical to the CPU path `clip_preprocess`, which
        is what the reference runs per image on the host: mm_utils.py:128-139 + CLIPImageProcessor). fp32 [n,3,S,S] on the device."""
        import numpy as np
        from ...engine import clip_preprocess_u8
        raw = [torch.from_numpy(np.ascontiguousarray(np.asarray(self.image_loader(p).convert("RGB"), dtype=np.uint8))) for p in image]
        return clip_preprocess_u8(raw, self.cfg.image_size, self.engine.device, pad=self.image_aspect_ratio == "pad")

    def _tokenize(self, questions: List[str], answers: List[str]):
        cache = self.__dict__.setdefault("_chunk_cache", {})
        if len(cache) > 65536:
            cache.clear()
        ids = [t5_tokenizer_image_token(q, self.tokenizer, chunk_cache=cache)[: self.context_len] for q in questions]
        labs = [t5_tokenizer_image_token(a, self.tokenizer, chunk_cache=cache)[: self.context_len] for a in answers]
        pad = getattr(self.tokenizer, "pad_token_id", 0) or 0
        # The splice kernel replaces exactly ONE image slot per prompt and embeds every other id: a caption that itself contains the
        # literal "<image>" would add a second slot (t5_tokenizer_image_token splits on it), and an id outside the vocabulary would index
        # past the embedding table. Refuse both on the host.
        for q, row in zip(questions, ids):
            if sum(1 for t in row if t == IMAGE_TOKEN_INDEX) != 1:
                raise ValueError(f"each prompt must contain exactly one {DEFAULT_IMAGE_TOKEN!r} placeholder (the caption must not contain it): {q!r}")
            if any((t < 0 and t != IMAGE_TOKEN_INDEX) or t >= self.cfg.vocab for t in row):
                raise ValueError("token id outside [0, vocab)")
        for row in labs:
            if any(t < 0 or t >= self.cfg.vocab for t in row):
                raise ValueError(f"answer token id outside [0, vocab) (answers must not contain {DEFAULT_IMAGE_TOKEN!r})")
        L, T = max(map(len, ids)), max(map(len, labs))
        if 0 < len(ids) <= getattr(self, "cuda_graph_max_pairs", 0):
            L = (L + 15) // 16 * 16          # few distinct shapes -> few graphs; padded columns are masked by `lens` in the engine
        input_ids = torch.full((len(ids), L), pad, dtype=torch.int32)
        labels = torch.full((len(labs), T), IGNORE_INDEX, dtype=torch.int32)
        lens = torch.zeros(len(ids), dtype=torch.int32)
        for i, (a, b) in enumerate(zip(ids, labs)):
            input_ids[i, : len(a)] = torch.tensor(a, dtype=torch.int32)
            labels[i, : len(b)] = torch.tensor(b, dtype=torch.int32)
            lens[i] = len(a)
        return input_ids, lens, labels

    # ---- scoring -------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, images: List[str], texts: List[str], question_template: str = default_question_template,
                answer_template: str = default_answer_template) -> torch.Tensor:
        assert len(images) == len(texts), "Number of images and texts must match"
        questions = [format_question(question_template.format(t), self.conversational_style) for t in texts]
        answers = [format_answer(answer_template.format(t), self.conversational_style) for t in texts]
        input_ids, lens, labels = self._tokenize(questions, answers)
        # identical paths are encoded once (Score.forward repeats one image N times, score.py:104-106)
        uniq: Dict[str, int] = {}
        index = []
        for p in images:
            index.append(uniq.setdefault(p, len(uniq)))
        pixels = self.load_images(list(uniq.keys()))
        dev = self.engine.device
        image_index = torch.tensor(index, dtype=torch.int32).to(dev, non_blocking=True)
        args = (pixels, input_ids.to(dev, non_blocking=True), lens.to(dev, non_blocking=True), labels.to(dev, non_blocking=True))
        kw = dict(image_index=image_index if len(uniq) != len(images) else None)
        if 0 < len(images) <= self.cuda_graph_max_pairs:
            try:
                return self.engine.score_tensors_graphed(*args, **kw).float().cpu()
            except RuntimeError as e:        # stream capture refused (e.g. another thread is issuing CUDA calls): same kernels, launched directly
                import warnings
                warnings.warn(f"CUDA-graph capture failed ({e}); small calls will be launched directly from now on")
                self.cuda_graph_max_pairs = 0
                torch.cuda.synchronize(dev)
        return self.engine.score_tensors(*args, **kw).float().cpu()
