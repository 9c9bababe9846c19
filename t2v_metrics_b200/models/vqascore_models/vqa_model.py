"""The VQAScore flavour of the plugin contract.

A VQAScore plugin scores a pair by asking the model a question built from the text and reading the probability of a fixed answer
(reference: t2v_metrics/models/vqascore_models/vqa_model.py:7-18). Both templates are ``str.format`` patterns with at most one ``{}``
slot that receives the pair's text -- e.g. 'Does this figure show "{}"? Please answer yes or no.' / 'Yes' (VQAScore), or '' / '{}'
(VisualGPTScore: the caption itself is the target sequence).
"""
from __future__ import annotations

import abc
from typing import List

import torch

from ..model import ScoreModel


class VQAScoreModel(ScoreModel):
    """Plugins return a CPU fp32 tensor with one probability-like score in [0, 1] per pair."""

    @abc.abstractmethod
    def forward(self, images: List[str], texts: List[str], question_template: str, answer_template: str) -> torch.Tensor:
        ...
