"""photoverse_b200 -- B200-native (sm_100a) implementation of PhotoVerse's dual-branch conditioning hot path.

Public surface mirrors the reference (idonahum/photoVerse):
    models/attention_processor.py -> photoverse_b200.attention_processor.PhotoVerseAttnProcessor2_0
    models/adapters.py            -> photoverse_b200.adapters.PhotoVerseAdapter
    models/unet.py                -> photoverse_b200.unet.{set_visual_cross_attention_adapter,
                                     get_visual_cross_attention_values_norm, set_cross_attention_layers_to_train}
    diffusers AttnProcessor2_0 on attn1 (models/unet.py:20-24) -> photoverse_b200.self_attention.SelfAttnProcessor (opt-in)
All arithmetic runs in libphotoverse_b200.so (hand-written CUDA, C ABI in include/photoverse_b200.h).
Importing the package does not load the library; the first op does, and raises if it was not built.
"""
from .adapters import PhotoVerseAdapter
from .attention_processor import PhotoVerseAttnProcessor, PhotoVerseAttnProcessor2_0
from .self_attention import SelfAttnProcessor, install_self_attention
from .unet import (get_visual_cross_attention_values_norm, set_cross_attention_layers_to_train,
                   set_visual_cross_attention_adapter)

__all__ = ["PhotoVerseAdapter", "PhotoVerseAttnProcessor", "PhotoVerseAttnProcessor2_0",
           "set_visual_cross_attention_adapter", "get_visual_cross_attention_values_norm",
           "set_cross_attention_layers_to_train", "SelfAttnProcessor", "install_self_attention"]
__version__ = "0.1.0"
