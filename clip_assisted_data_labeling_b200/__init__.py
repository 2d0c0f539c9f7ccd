"""B200-native (sm_100a) drop-in for the data-parallel hot path of aiXander/CLIP_assisted_data_labeling:
the 4-crop CLIP-ViT embedding pass (reference utils/embedder.py, _1_embed_with_CLIP.py) and the
all-to-all cosine duplicate search (reference _2_remove_duplicates.py).

Importing this package does not touch CUDA or load libb2c.so (DataLoader workers re-import it under
``spawn``); the library is loaded on first use and there is no fallback when it is missing.
"""
from .vit_arch import ARCHS, CROP_NAMES, OPENAI_MEAN, OPENAI_STD  # noqa: F401

__all__ = ["ARCHS", "CROP_NAMES", "OPENAI_MEAN", "OPENAI_STD"]
