"""Plugin base class -- same contract as the reference's ScoreModel (t2v_metrics/models/model.py:10-47)."""
from abc import ABC, abstractmethod
from typing import List
import os

import numpy as np
import torch
from PIL import Image

from ..constants import HF_CACHE_DIR


def image_loader(image_path):
    """model.py:10-14: .npy arrays are BGR -> RGB, everything else through PIL .convert('RGB')."""
    if image_path.split('.')[-1] == 'npy':
        return Image.fromarray(np.load(image_path)[:, :, [2, 1, 0]], 'RGB')
    return Image.open(image_path).convert("RGB")


class ScoreModel(ABC):
    def __init__(self, model_name='clip-flant5-xxl', device='cuda', cache_dir=HF_CACHE_DIR):
        self.model_name = model_name
        self.device = device
        self.cache_dir = cache_dir
        if self.cache_dir and not os.path.exists(self.cache_dir):
            os.makedirs(self.cache_dir, exist_ok=True)
        self.image_loader = image_loader
        self.load_model()

    @abstractmethod
    def load_model(self):
        """Load the model, tokenizer, and etc."""

    @abstractmethod
    def load_images(self, image: List[str]) -> torch.Tensor:
        """Load the image(s), and return a tensor (after preprocessing) put on self.device"""

    @abstractmethod
    def forward(self, images: List[str], texts: List[str]) -> torch.Tensor:
        """Return n scores for n (image, text) pairs"""
