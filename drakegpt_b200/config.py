"""Paths and hyper-parameters -- same names and values as the reference's src/config.py:1-44.

Paths are resolved relative to the repository root instead of the process cwd (the reference
must be run from inside src/); override with the DRAKEGPT_ROOT environment variable.
"""
import os
from pathlib import Path

_ROOT = Path(os.environ.get("DRAKEGPT_ROOT", Path(__file__).resolve().parent.parent))

MODEL_DIR = _ROOT / "model"
DATA_DIR = _ROOT / "data"
INFERENCE_DIR = _ROOT / "inference"

DATA = {
    "input": DATA_DIR / "Drake_lyrics.txt",
    "train": DATA_DIR / "train_data.pt",
    "val": DATA_DIR / "val_data.pt",
    "drake": DATA_DIR / "drake.csv",
}

PARAMS = {
    "context_length": 8,
    "batch_size": 32,
    "base_lr": 1e-3,
    "max_lr": 5e-3,
    "betas": (0.9, 0.95),
    "embedding_dim": 32,
    "head_size": 32,
    "num_heads": 4,
    "num_layers": 3,
    "dropout": 0.1,
}

SCALE_PARAMS = {
    "context_length": 256,
    "batch_size": 64,
    "base_lr": 3e-4,
    "max_lr": 6e-4,
    "betas": (0.9, 0.95),
    "embedding_dim": 384,
    "head_size": 64,
    "num_heads": 6,
    "num_layers": 6,
    "dropout": 0.2,
}

TRAIN = {
    "iters": 10000,
    "eval_iters": 200,
    "eval_interval": 500,
}
