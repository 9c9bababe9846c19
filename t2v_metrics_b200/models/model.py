"""Base class of every scoring plugin.

The reference's plugin contract (t2v_metrics/models/model.py:10-47) is small and the callers above it rely on exactly these pieces:
`Score.prepare_scoremodel` constructs a plugin with (model_name, device, cache_dir), the constructor must finish with the model loaded,
and `Score.forward` then calls `plugin.forward(images, texts, **kwargs)` with file paths. This module restates that contract for the
engine-backed plugins; nothing here touches the GPU.
"""
from __future__ import annotations

import abc
import pathlib
from typing import Callable, List

import numpy as np
import torch
from PIL import Image

from ..constants import HF_CACHE_DIR


def image_loader(image_path: str) -> Image.Image:
    """Decode one image file to an RGB PIL image.

    Same two cases as the reference loader (model.py:10-14): a ``.npy`` file holds an OpenCV-style BGR array whose channels are
    reversed, anything else is whatever PIL can open, converted to RGB (drops alpha, expands greyscale / palette images).
    """
    path = pathlib.Path(image_path)
    if path.suffix.lower() == ".npy":
        bgr = np.load(path)
        return Image.fromarray(np.ascontiguousarray(bgr[..., ::-1]), "RGB")
    with Image.open(path) as im:
        return im.convert("RGB")


class ScoreModel(abc.ABC):
    """A plugin owns one loaded model and turns (image path, text) pairs into scores.

    Attributes the callers read: ``model_name``, ``device``, ``cache_dir``, ``image_loader`` (overridable decoder hook).
    Subclasses implement :meth:`load_model` (called once, at the end of the constructor), :meth:`load_images` and :meth:`forward`.
    """

    def __init__(self, model_name: str = "clip-flant5-xxl", device: str = "cuda", cache_dir: str = HF_CACHE_DIR):
        self.model_name, self.device, self.cache_dir = model_name, device, cache_dir
        if cache_dir:
            pathlib.Path(cache_dir).mkdir(parents=True, exist_ok=True)
        self.image_loader: Callable[[str], Image.Image] = image_loader
        self.load_model()

    @abc.abstractmethod
    def load_model(self) -> None:
        """Bring up tokenizer / processor / weights; after this call the plugin must be ready to score."""

    @abc.abstractmethod
    def load_images(self, image: List[str]) -> torch.Tensor:
        """Decode and pre-process the given files; the result lives on the plugin's device."""

    @abc.abstractmethod
    def forward(self, images: List[str], texts: List[str]) -> torch.Tensor:
        """One score per (images[i], texts[i]) pair, as a 1-D tensor of len(images) elements."""
