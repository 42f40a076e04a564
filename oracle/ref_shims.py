"""Import the REFERENCE's own hot-path modules from /root/reference (TEST INFRASTRUCTURE).

Used by ``oracle/gen_golden.py`` (and by tests that run only when /root/reference is present) to execute the
reference itself: ``APadapter/ap_adapter/attention_processor.py`` imports cleanly; ``audio_encoder/*.py`` need
``sys.modules`` shims for ``timm`` (restated Block / to_2tuple), ``librosa`` and ``matplotlib``, none of which
is installed here.  Nothing from the reference is copied — the modules are loaded from where they lie (on a box without the
tree: from the bytecode ``oracle/build_ref.py`` compiled out of it).
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("APX_REFERENCE_ROOT", "/root/reference")


REF_BUILT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/build_ref.py compiles the processor module into here


def reference_available() -> bool:
    """The whole reference tree is mounted (this container; never the GPU box)."""
    return os.path.isfile(os.path.join(REF_ROOT, "APadapter", "ap_adapter", "attention_processor.py"))


def _built_module_ok(path: str) -> bool:
    """oracle/_ref/attention_processor.pyc exists and was compiled by this interpreter's bytecode version."""
    try:
        with open(path, "rb") as f:
            return f.read(4) == importlib.util.MAGIC_NUMBER
    except OSError:
        return False


def processor_module_path() -> str | None:
    """The reference's attention_processor: the source in the mounted tree, else the module compiled from it under oracle/_ref."""
    src = os.path.join(REF_ROOT, "APadapter", "ap_adapter", "attention_processor.py")
    if os.path.isfile(src):
        return src
    built = os.path.join(REF_BUILT, "attention_processor.pyc")
    return built if _built_module_ok(built) else None


def reference_processors_available() -> bool:
    return processor_module_path() is not None


def _load(name: str, path: str):
    loader = importlib.machinery.SourcelessFileLoader(name, path) if path.endswith(".pyc") else None
    spec = importlib.util.spec_from_file_location(name, path, loader=loader)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_attention_processor():
    """The reference's attention_processor module (IPAttnProcessor2_0, AttnProcessor2_0, ...)."""
    path = processor_module_path()
    if path is None:
        raise FileNotFoundError("neither /root/reference nor oracle/_ref holds the attention_processor module (run oracle/build_ref.py)")
    return sys.modules.get("apx_ref_attention_processor") or _load("apx_ref_attention_processor", path)


class StubAttention(nn.Module):
    """Duck-typed stand-in for diffusers' ``Attention`` exposing exactly what the reference processors read
    (SURVEY.md §8b): heads, to_q/to_k/to_v, to_out = [Linear, Dropout], the norm hooks (all off for
    ``BasicTransformerBlock.attn2``), ``prepare_attention_mask``, ``residual_connection``, ``rescale_output_factor``."""

    def __init__(self, query_dim: int, cross_attention_dim: int | None, heads: int, bias: bool = False):
        super().__init__()
        ctx = cross_attention_dim or query_dim
        self.heads = heads
        self.to_q = nn.Linear(query_dim, query_dim, bias=bias)
        self.to_k = nn.Linear(ctx, query_dim, bias=bias)
        self.to_v = nn.Linear(ctx, query_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(query_dim, query_dim, bias=True), nn.Dropout(0.0)])
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0

    def prepare_attention_mask(self, attention_mask, target_length, batch_size):
        if attention_mask is None:
            return None
        if attention_mask.shape[-1] != target_length:
            attention_mask = torch.nn.functional.pad(attention_mask, (0, target_length), value=0.0)
        if attention_mask.shape[0] < batch_size * self.heads:
            attention_mask = attention_mask.repeat_interleave(self.heads, dim=0)
        return attention_mask


# ---- timm / librosa / matplotlib shims so audio_encoder/*.py import unchanged -----------------------
class _TimmAttention(nn.Module):
    def __init__(self, dim, num_heads, qkv_bias=True):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        b, n, c = x.shape
        qkv = self.qkv(x).reshape(b, n, 3, self.num_heads, c // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = ((q * self.scale) @ k.transpose(-2, -1)).softmax(dim=-1)
        return self.proj((attn @ v).transpose(1, 2).reshape(b, n, c))


class _TimmMlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _TimmBlock(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, norm_layer=nn.LayerNorm, **_):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = _TimmAttention(dim, num_heads, qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = _TimmMlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


def _install_shims():
    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)  # noqa: E731
    vt = mod("timm.models.vision_transformer", Block=_TimmBlock, VisionTransformer=nn.Module)
    layers = mod("timm.models.layers", to_2tuple=to_2tuple)
    models = mod("timm.models", vision_transformer=vt, layers=layers)
    mod("timm", models=models)
    mod("librosa.display")
    mod("librosa", display=sys.modules["librosa.display"])
    mod("matplotlib.pyplot")
    mod("matplotlib", pyplot=sys.modules["matplotlib.pyplot"])


def load_audio_encoder():
    """Returns (models_mae, AudioMAE) modules of the reference, imported as the package ``apx_ref_audio_encoder``."""
    _install_shims()
    pkg_dir = os.path.join(REF_ROOT, "audio_encoder")
    pkg = types.ModuleType("apx_ref_audio_encoder")
    pkg.__path__ = [pkg_dir]
    sys.modules["apx_ref_audio_encoder"] = pkg
    models_vit = _load("apx_ref_audio_encoder.models_vit", os.path.join(pkg_dir, "models_vit.py"))
    models_mae = _load("apx_ref_audio_encoder.models_mae", os.path.join(pkg_dir, "models_mae.py"))
    pkg.models_vit, pkg.models_mae = models_vit, models_mae
    audiomae = _load("apx_ref_audio_encoder.AudioMAE", os.path.join(pkg_dir, "AudioMAE.py"))
    return models_mae, audiomae
