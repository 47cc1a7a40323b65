"""CPU: the C-ABI shared library loads and exports every symbol include/photoverse_b200.h declares; without a
GPU the compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "photoverse_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pv_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path_entry_points():
    syms = declared_symbols()
    for s in ("pv_version", "pv_last_error", "pv_launch_count", "pv_pack_weight", "pv_linear_fwd", "pv_kv_pack_fwd",
              "pv_dual_attn_fwd", "pv_ln_lrelu_fwd", "pv_group_mean_fwd"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from photoverse_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build with `python -m photoverse_b200.build`"
    l = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(l, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert _lib.lib().pv_version() == 2


def test_python_binding_covers_the_header():
    from photoverse_b200 import _lib
    assert set(declared_symbols()) <= set(_lib._SIGNATURES), set(declared_symbols()) - set(_lib._SIGNATURES)


def test_invalid_arguments_set_an_error_message():
    from photoverse_b200 import _lib
    l = _lib.lib()
    rc = l.pv_set_option(b"no_such_option", 1)
    assert rc == 1 and b"no_such_option" in l.pv_last_error()
    rc = l.pv_linear_fwd(1, 1, None, None, None, None, 1, 1, 8, 1, 8, 8, 8, 0, 0, 0, 0, None)
    assert rc == 1 and b"null" in l.pv_last_error()
    assert l.pv_kv_tile_bytes(1, 320, 8, 77, 5) == 96 * 48 * 2
    assert l.pv_kv_tile_bytes(1, 640, 8, 77, 5) == 96 * 80 * 2
    assert l.pv_kv_tile_bytes(0, 1280, 8, 77, 5) == 82 * 160 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_a_gpu():
    """Product modules refuse CPU tensors instead of silently computing elsewhere."""
    from photoverse_b200 import PhotoVerseAdapter, PhotoVerseAttnProcessor2_0
    from photoverse_b200.host.unet_sd15 import Attention
    attn = Attention(320, 768, 8)
    proc = PhotoVerseAttnProcessor2_0(hidden_size=320, cross_attention_dim=768)
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        proc(attn, torch.zeros(1, 128, 320), encoder_hidden_states=(torch.zeros(1, 77, 768), torch.zeros(1, 5, 768)))
    ad = PhotoVerseAdapter(num_tokens=1)
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        ad([torch.zeros(1, 257, 1024)])
