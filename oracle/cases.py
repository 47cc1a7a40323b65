"""TEST INFRASTRUCTURE ONLY -- the golden-vector case table shared by make_golden.py and tests/.

Inputs and weights are regenerated from ``detgen`` (platform independent); only outputs are stored
under tests/golden/.  Shapes are the SD-1.5 attn2 families of SURVEY.md §2.2 at reduced S / B so the
fixtures stay small, plus the reference's edge cases: Li=1 (generate path, token_index=0,
SURVEY §0.1 D5), Li=16 (microbench max), a ragged S that is not a multiple of the 128-row tile,
LoRA on q/k/v, and the two grad-mode fusion branches (attention_processor.py:413-420).
"""
from dataclasses import dataclass
from typing import Optional

import torch

from . import detgen
from .processor_oracle import LoraWeights, ProcessorWeights


@dataclass(frozen=True)
class ProcCase:
    name: str
    B: int
    S: int
    C: int
    Li: int
    Lt: int = 77
    H: int = 8
    Dc: int = 768
    lora_r: int = 0
    w_text: float = 1.0
    w_img: float = 1.0
    seed: int = 0


PROC_CASES = [
    ProcCase("c320_s256_li5", B=2, S=256, C=320, Li=5, seed=11),
    ProcCase("c640_s128_li1", B=2, S=128, C=640, Li=1, seed=12),
    ProcCase("c1280_s64_li16", B=1, S=64, C=1280, Li=16, seed=13),
    ProcCase("c320_s200_ragged_li4", B=1, S=200, C=320, Li=4, seed=14),
    ProcCase("c320_s128_lora4", B=1, S=128, C=320, Li=5, lora_r=4, seed=15),
    ProcCase("c640_s128_lora16_textonly", B=1, S=128, C=640, Li=5, lora_r=16, w_text=2.0, w_img=0.0, seed=16),
    ProcCase("c320_s128_imgonly", B=2, S=128, C=320, Li=5, w_text=0.0, w_img=2.0, seed=17),
    ProcCase("c1280_s128_li5", B=1, S=128, C=1280, Li=5, seed=18),
]


def proc_weights(c: ProcCase, dtype=torch.float32) -> ProcessorWeights:
    s = c.seed * 100
    w = ProcessorWeights(
        to_q=detgen.uniform_linear(c.C, c.C, s + 1, dtype),
        to_k=detgen.uniform_linear(c.C, c.Dc, s + 2, dtype),
        to_v=detgen.uniform_linear(c.C, c.Dc, s + 3, dtype),
        to_out_w=detgen.uniform_linear(c.C, c.C, s + 4, dtype),
        to_out_b=detgen.uniform_bias(c.C, c.C, s + 5, dtype),
        to_k_ip=detgen.uniform_linear(c.C, c.Dc, s + 6, dtype),
        to_v_ip=detgen.uniform_linear(c.C, c.Dc, s + 7, dtype),
        heads=c.H,
    )
    if c.lora_r:
        r = c.lora_r
        # peft initialises B to zeros (exact no-op); use non-zero B so the LoRA path is exercised.
        for j, (name, fin) in enumerate((("to_q", c.C), ("to_k", c.Dc), ("to_v", c.Dc))):
            w.lora[name] = LoraWeights(
                A=detgen.uniform_linear(r, fin, s + 20 + 2 * j, dtype),
                B=detgen.uniform((c.C, r), s + 21 + 2 * j, -0.2, 0.2, dtype),
                scaling=1.0 / r,
            )
    return w


def proc_inputs(c: ProcCase, dtype=torch.float32):
    s = c.seed * 100
    x = detgen.unit_variance((c.B, c.S, c.C), s + 50, dtype)
    text = detgen.unit_variance((c.B, c.Lt, c.Dc), s + 51, dtype)
    img = detgen.unit_variance((c.B, c.Li, c.Dc), s + 52, dtype)
    return x, text, img


@dataclass(frozen=True)
class AdapterCase:
    name: str
    B: int
    T: int                      # number of token heads the adapter owns
    token_index: Optional[object]
    tokens: int = 257
    seed: int = 0


ADAPTER_CASES = [
    AdapterCase("adapter_full_t5", B=2, T=5, token_index=None, seed=31),
    AdapterCase("adapter_idx0_t5", B=2, T=5, token_index=0, seed=31),
    AdapterCase("adapter_idx3_t5", B=1, T=5, token_index=3, seed=31),
    AdapterCase("adapter_fullstr_t2", B=3, T=2, token_index="full", seed=32),
]


def adapter_inputs(c: AdapterCase, dtype=torch.float32):
    return [detgen.unit_variance((c.B, c.tokens, 1024), c.seed * 100 + 70 + i, dtype) for i in range(c.T)]
