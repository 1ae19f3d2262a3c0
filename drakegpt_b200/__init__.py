"""drakegpt_b200 -- B200-native (sm_100a) training step and generation for the DrakeGPT
character-level language models, behind the reference's own PyTorch module API.

    from drakegpt_b200.model import TransformerLM
    from drakegpt_b200.model_component import Head, MultiHeadAttention, FeedForward, Block, ResidualBlock

Compute runs in hand-written CUDA kernels loaded from ``csrc/libdrakegpt_b200.so``
(C-ABI in ``include/drakegpt_b200.h``).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
__all__ = ["model", "model_component", "ops", "engine", "optim", "config", "preprocessing", "train", "inference"]
