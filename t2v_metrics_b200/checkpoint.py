"""Local checkpoint loading for the engine-backed plugins.

The reference loads HuggingFace repositories with `from_pretrained` (t2v_metrics/models/vqascore_models/mm_utils.py:182-241 for
CLIP-FlanT5, qwen2vl_model.py:110-133 for Qwen2.5-VL): sharded `*.safetensors` / `pytorch_model-*.bin` files listed by an
`*.index.json`, with on-disk tensor names that transformers remaps while loading. There is no network here, so the plugins take a local
path (`checkpoint=`); this module reads what such a path can be -- one file or a repository directory -- and normalises the names to
the in-memory HF names `engine.convert_state_dict` / `engine.convert_qwen_state_dict` consume. Host-side only, no GPU.
"""
from __future__ import annotations

import json
import os
import re
from typing import Dict, Iterable, Optional

import torch

_INDEX_FILES = ("model.safetensors.index.json", "pytorch_model.bin.index.json")
_SINGLE_FILES = ("model.safetensors", "pytorch_model.bin", "consolidated.safetensors")


def _load_file(path: str) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    blob = torch.load(path, map_location="cpu", weights_only=True)
    if isinstance(blob, dict) and "state_dict" in blob and isinstance(blob["state_dict"], dict):
        blob = blob["state_dict"]
    return blob


def load_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """One `.safetensors` / `.bin` / `.pt` file, or an HF repository directory: an index json + its shards, a single weights file, or
    (last resort) every `*.safetensors` in the directory. Raises FileNotFoundError with the offline hint otherwise."""
    if os.path.isfile(path):
        return _load_file(path)
    if os.path.isdir(path):
        for name in _INDEX_FILES:
            idx = os.path.join(path, name)
            if os.path.isfile(idx):
                with open(idx) as f:
                    weight_map = json.load(f)["weight_map"]
                out: Dict[str, torch.Tensor] = {}
                for shard in sorted(set(weight_map.values())):
                    part = _load_file(os.path.join(path, shard))
                    out.update({k: v for k, v in part.items() if weight_map.get(k) == shard})
                missing = set(weight_map) - set(out)
                if missing:
                    raise KeyError(f"{idx} lists tensors that no shard contains: {sorted(missing)[:5]} ...")
                return out
        for name in _SINGLE_FILES:
            f = os.path.join(path, name)
            if os.path.isfile(f):
                return _load_file(f)
        shards = sorted(f for f in os.listdir(path) if f.endswith(".safetensors"))
        if shards:
            out = {}
            for f in shards:
                out.update(_load_file(os.path.join(path, f)))
            return out
    raise FileNotFoundError(
        f"no weights at {path!r}: pass `checkpoint=` pointing at a local .safetensors/.bin/.pt file or at a downloaded HF repository "
        "directory (index json + shards), or pass `state_dict=`. This build has no network access.")


def generation_config_value(path: str, key: str, default):
    """A field of the generation_config.json stored with the checkpoint (what from_pretrained attaches to model.generation_config)."""
    d = path if os.path.isdir(path) else os.path.dirname(path)
    f = os.path.join(d, "generation_config.json") if d else ""
    if f and os.path.isfile(f):
        with open(f) as fh:
            v = json.load(fh).get(key, default)
        return default if v is None else v
    return default


# ------------------------------------------------------------------------------------------------ Qwen2.5-VL
def normalise_qwen_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Published Qwen2.5-VL checkpoints store `visual.*`, `model.layers.*`, `model.embed_tokens.*`, `model.norm.*`, `lm_head.*`;
    transformers >= 4.52 holds them in memory as `model.visual.*`, `model.language_model.*`, `lm_head.*` (its conversion mapping applies
    `^visual -> model.visual` and `^model(?!\\.(language_model|visual)) -> model.language_model`). Accept either; return the in-memory
    names."""
    out = {}
    for k, v in sd.items():
        if k.startswith("visual."):
            k = "model." + k
        elif k.startswith("model.") and not k.startswith(("model.visual.", "model.language_model.")):
            k = "model.language_model." + k[len("model."):]
        out[k] = v
    return out


# ------------------------------------------------------------------------------------------------ CLIP-FlanT5
_T5_ROOTS = ("shared.", "encoder.", "decoder.", "lm_head.")


def normalise_clipt5_keys(sd: Dict[str, torch.Tensor], vision_sd: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """-> `vision_tower.vision_model.*`, `mm_projector.{0,2}.*`, and plain T5ForConditionalGeneration names.

    The LLaVA-style CLIP-FlanT5 checkpoints nest the tower and the projector under wrapper attributes (`...vision_tower.vision_tower.
    vision_model.*`, `...mm_projector.*`, possibly behind a `model.` prefix) and the loader brings the CLIP tower in SEPARATELY when the
    checkpoint does not carry it (`model.get_vision_tower().load_model()`, mm_utils.py:226-227): pass that second state dict
    (`openai/clip-vit-large-patch14-336`, names `vision_model.*`) as `vision_sd`."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        m = re.search(r"vision_model\.(.*)$", k)
        if m and "vision_tower" in k:
            out["vision_tower.vision_model." + m.group(1)] = v
            continue
        m = re.search(r"mm_projector\.(\d+\.(?:weight|bias))$", k)
        if m:
            out["mm_projector." + m.group(1)] = v
            continue
        for root in _T5_ROOTS:
            i = k.find(root)
            if i == 0 or (i > 0 and k[i - 1] == "."):
                if not any(k[:i].endswith(p) for p in ("vision_tower.", "vision_model.")):
                    out[k[i:]] = v
                break
    if vision_sd is not None:
        for k, v in vision_sd.items():
            m = re.search(r"vision_model\.(.*)$", k)
            if m:
                out.setdefault("vision_tower.vision_model." + m.group(1), v)
    if not any(k.startswith("vision_tower.vision_model.") for k in out):
        raise KeyError("the checkpoint holds no CLIP vision tower (`...vision_tower...vision_model.*`): the reference loads it separately from "
                       "openai/clip-vit-large-patch14-336 (mm_utils.py:226-227) -- pass `vision_tower_checkpoint=` with a local copy")
    if "encoder.embed_tokens.weight" in out and "shared.weight" not in out:
        out["shared.weight"] = out["encoder.embed_tokens.weight"]
    return out
