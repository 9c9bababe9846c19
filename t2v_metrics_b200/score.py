"""Score orchestration above the plugin boundary -- same call surface as the reference's t2v_metrics/score.py
(`Score.forward` :47-113 -> [M, N] tensor on `device`; `Score.batch_forward` :115-156 -> [samples, visuals, texts]).

Differences, all above the drop-in boundary and result-preserving:
  * `forward` hands ALL M x N pairs to the plugin in one call (chunked by `max_pairs`); the plugin encodes each distinct
    image once. The reference loops `self.model.forward([image] * N, texts)` per image (score.py:104-106).
  * `batch_forward` scores a whole DataLoader batch per (visual, text) slot, as v3.0 did, instead of one pair at a time
    (v3.1 score.py:143-153).
  * video inputs (the cv2 concat path, score.py:72-98) are CPU pre-processing outside this engine's scope and raise.
"""
from typing import List, Optional, TypedDict, Union

import torch
import torch.nn as nn

from .constants import HF_CACHE_DIR

VIDEO_EXTENSIONS = {'.mp4', '.avi', '.mov', '.mkv'}


class ImageTextDict(TypedDict):
    images: List[str]
    texts: List[str]


class Score(nn.Module):
    def __init__(self, model: str, device: str = 'cuda', cache_dir: str = HF_CACHE_DIR, **kwargs):
        super().__init__()
        assert model in self.list_all_models()
        self.device = device
        self.model = self.prepare_scoremodel(model, device, cache_dir, **kwargs)
        self.model_name = model
        self.max_pairs = 64

    def prepare_scoremodel(self, model: str, device: str, cache_dir: str, **kwargs):
        raise NotImplementedError("Subclasses must implement prepare_scoremodel")

    def list_all_models(self) -> List[str]:
        raise NotImplementedError("Subclasses must implement list_all_models")

    def forward(self, images: Optional[Union[str, List[str]]] = None, texts: Optional[Union[str, List[str]]] = None,
                num_frames: Optional[int] = 8, **kwargs) -> torch.Tensor:
        """m images x n texts -> [m, n] tensor of scores on self.device."""
        if isinstance(images, str):
            images = [images]
        if isinstance(texts, str):
            texts = [texts]
        if any(isinstance(img, str) and img[-4:].lower() in VIDEO_EXTENSIONS for img in images):
            raise NotImplementedError("video inputs (frame extraction + concat, reference score.py:72-98) are outside the "
                                      "B200 engine's hot-path scope; pass image files")
        m, n = len(images), len(texts)
        pair_images = [img for img in images for _ in range(n)]
        pair_texts = [t for _ in range(m) for t in texts]
        out = []
        for s in range(0, m * n, self.max_pairs):
            out.append(self.model.forward(pair_images[s:s + self.max_pairs], pair_texts[s:s + self.max_pairs], **kwargs))
        return torch.cat(out).view(m, n).to(self.device)

    def batch_forward(self, dataset: List[ImageTextDict], batch_size: int = 16, num_frames: int = 4, **kwargs) -> torch.Tensor:
        """[num_samples, num_visuals, num_texts] scores for a dataset of {'images': [...], 'texts': [...]} items."""
        from torch.utils.data import DataLoader
        num_samples = len(dataset)
        if "videos" in dataset[0]:
            raise NotImplementedError("video datasets are outside the B200 engine's hot-path scope")
        num_visuals = len(dataset[0]['images'])
        num_texts = len(dataset[0]['texts'])
        scores = torch.zeros(num_samples, num_visuals, num_texts).to(self.device)
        dataloader = DataLoader(dataset, batch_size=batch_size, shuffle=False)
        counter = 0
        for batch_idx, batch in enumerate(dataloader):
            cur = len(batch['images'][0])
            assert len(batch['images']) == num_visuals and len(batch['texts']) == num_texts
            for vis_idx in range(num_visuals):
                for text_idx in range(num_texts):
                    scores[counter:counter + cur, vis_idx, text_idx] = self.model.forward(
                        list(batch['images'][vis_idx]), list(batch['texts'][text_idx]), **kwargs).to(self.device)
            counter += cur
        return scores
