"""Host graph of the AudioLDM2 UNet: the CALLER of the adapter hot path, restated in plain PyTorch.

The reference builds its UNet from ``pipeline/modeling_audioldm2.py`` (wiring) on top of diffusers 0.21.2 leaf modules
(``Attention``, ``BasicTransformerBlock``, ``Transformer2DModel``, ``ResnetBlock2D``, ``Down/Upsample2D``, ``Timesteps``,
``TimestepEmbedding``; imported at modeling_audioldm2.py:22-41).  diffusers is not installed on the B200 image and the
hub config / weights are unreachable, so this module restates

* the leaf semantics with diffusers' parameter names, so a real ``unet`` state dict maps 1:1;
* the reference wiring: 4 ``Transformer2DModel`` per resnet layer with per-slot context routing
  (modeling_audioldm2.py:1140-1156, 1317-1333, 1492-1508), the processor registry
  ``attn_processors`` / ``set_attn_processor`` (:517-573) and ``forward`` (:663-873).

Nothing here is a kernel target: it is the host code around the 32 ``attn2`` adapter sites, whose arithmetic runs in
``IPAttnProcessor2_0`` -> libapx.so.  Geometry defaults are the ones the shipped adapter weights imply (SURVEY.md §8
a-4): ``block_out_channels=(128,256,384,640)``, 2 resnet layers, 2 transformer layers, per-slot cross dims
``(None, 768, 1024, None)``, 8 heads, 8 latent channels.
"""
from __future__ import annotations

import math
import os
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .attention_processor import AttnProcessor2_0


# ---------------------------------------------------------------------------------------------------------
# channels-last inference route (the neighbours of the adapter sites)
# ---------------------------------------------------------------------------------------------------------
class NHWC:
    """Inference on CUDA in 16 bit runs the convolutional trunk in channels-last memory with libapx's fused GroupNorm kernel:

    * ``ResnetBlock2D``: GroupNorm+SiLU is one kernel (``apx_groupnorm_nhwc``) that also absorbs conv1's bias and the time-embedding
      add in front of norm2; shortcut + conv2 + both conv biases are one kernel (``apx_add_bias_nhwc``); cuDNN gets NHWC tensors and
      pre-converted weights, so its NCHW<->NHWC transposes disappear;
    * ``Transformer2DModel``: GroupNorm writes NHWC, whose (B, H*W, C) token view is free; the 1x1 proj_in / proj_out convolutions
      are linears on that view (bias in the cuBLAS epilogue), and the result goes back as a channels-last view;
    * the 22 time-embedding projections of a forward are ONE GEMM.

    Same arithmetic as the stock graph up to 16-bit rounding points (fewer of them).  ``APX_NHWC=0`` / ``NHWC.enabled = False`` keeps
    the stock NCHW graph (what the `torch_gpu` bench leg and the parity tests compare against)."""

    enabled = os.environ.get("APX_NHWC", "1") != "0"
    active = False          # set by AudioLDM2UNet2DConditionModel.forward for the duration of one forward

    @staticmethod
    def usable(x: torch.Tensor) -> bool:
        return (NHWC.enabled and x.is_cuda and x.dtype in (torch.float16, torch.bfloat16) and not torch.is_grad_enabled()
                and not x.requires_grad)


_CACHE_ATTR = "_apx_cache"


def _cached(module: nn.Module, tag: str, params, build):
    """Per-module cache of a derived tensor (fp32 copy, channels-last copy, concatenation ...) keyed on the identity and version of
    the source parameters, which are kept referenced so their addresses cannot be recycled under the key."""
    cache = module.__dict__.setdefault(_CACHE_ATTR, {})
    key = tuple((p.data_ptr(), p._version, p.dtype, p.device) for p in params)
    hit = cache.get(tag)
    if hit is not None and hit[0] == key and all(a is b for a, b in zip(hit[2], params)):
        return hit[1]
    with torch.no_grad():
        val = build()
    cache[tag] = (key, val, tuple(params))
    return val


def _f32(module, name):
    p = getattr(module, name)
    return _cached(module, "f32:" + name, (p,), lambda: p.detach().float().contiguous())


def _cl_weight(conv: nn.Conv2d):
    return _cached(conv, "clw", (conv.weight,), lambda: conv.weight.detach().contiguous(memory_format=torch.channels_last))


def _as_cl(x: torch.Tensor) -> torch.Tensor:
    return x if x.is_contiguous(memory_format=torch.channels_last) else x.contiguous(memory_format=torch.channels_last)


def _gn_nhwc(x, norm: nn.GroupNorm, silu: bool, chan_bias=None, bc_add=None):
    """[SiLU] GroupNorm(x + chan_bias + bc_add) on a channels-last tensor: libapx where the shape is covered, torch ops otherwise."""
    x = _as_cl(x)
    b, c, h, w = x.shape
    if _lib.groupnorm_nhwc_ok(h * w, c, norm.num_groups):
        return _lib.groupnorm_nhwc(x, _f32(norm, "weight"), _f32(norm, "bias"), norm.num_groups, norm.eps, silu, chan_bias, bc_add)
    if chan_bias is not None:
        x = x + chan_bias.to(x.dtype)[None, :, None, None]
    if bc_add is not None:
        x = x + bc_add[:, :, None, None]
    y = norm(x)
    return F.silu(y) if silu else y


# ---------------------------------------------------------------------------------------------------------
# diffusers leaf modules (restated)
# ---------------------------------------------------------------------------------------------------------
class Attention(nn.Module):
    """diffusers ``Attention`` as the processors see it (SURVEY.md §8b "What it reads from attn")."""

    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64,
                 dropout: float = 0.0, bias: bool = False, processor: Optional[nn.Module] = None):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.cross_attention_dim = cross_attention_dim or query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, inner, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(dropout)])
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.set_processor(processor if processor is not None else AttnProcessor2_0())

    def set_processor(self, processor) -> None:
        # a Module processor is registered as a sub-module (so AttnProcsLayers / .to() see its parameters);
        # replacing it by a non-Module drops the registration
        if hasattr(self, "processor") and isinstance(self.processor, nn.Module) and not isinstance(processor, nn.Module):
            self._modules.pop("processor")
        self.processor = processor

    def get_processor(self):
        return self.processor

    def prepare_attention_mask(self, attention_mask, target_length: int, batch_size: int):
        if attention_mask is None:
            return None
        if attention_mask.shape[-1] != target_length:
            # pads BY target_length (diffusers 0.21 behaviour, the origin of attention_processor.py:424-428)
            attention_mask = F.pad(attention_mask, (0, target_length), value=0.0)
        if attention_mask.shape[0] < batch_size * self.heads:
            attention_mask = attention_mask.repeat_interleave(self.heads, dim=0)
        return attention_mask

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    """diffusers FeedForward(dim, activation_fn="geglu"): Linear(dim, 8 dim) -> hidden * gelu(gate) -> Linear(4 dim, dim).

    `fused()` is the inference route of the restated host graph: two tcgen05 GEMMs, the first with the GEGLU in its
    epilogue (projection rows interleaved once so a (hidden_i, gate_i) pair shares an accumulator tile), the second with
    bias + the block's residual add in its epilogue.  State-dict keys are the stock ones (net.0.proj.*, net.2.*)."""

    # class-wide switch (the parity tests turn it off to get the unfused torch graph; APX_FF_FUSED=0 does the same for A/B runs)
    use_fused = os.environ.get("APX_FF_FUSED", "1") != "0"

    def __init__(self, dim: int, mult: int = 4, dropout: float = 0.0):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(dropout), nn.Linear(dim * mult, dim)])
        self._fused = None

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x

    def fused_ok(self, x) -> bool:
        return (self.use_fused and x.is_cuda and x.dtype in (torch.float16, torch.bfloat16) and not torch.is_grad_enabled()
                and self.net[0].proj.weight.dtype == x.dtype and x.shape[-1] % 64 == 0)

    def _packed(self):
        proj, out = self.net[0].proj, self.net[2]
        # keyed on the source tensors themselves (kept referenced, so an address cannot be recycled under the cache) + versions
        src = (proj.weight, proj.bias, out.bias)
        key = tuple((t._version, t.data_ptr(), t.dtype) for t in src)
        c = self._fused
        if c is None or c[0] != key or any(a is not b for a, b in zip(c[4], src)):
            inner = proj.weight.shape[0] // 2
            with torch.no_grad():
                w = torch.stack((proj.weight[:inner], proj.weight[inner:]), dim=1).reshape(2 * inner, -1).contiguous()
                b = torch.stack((proj.bias[:inner], proj.bias[inner:]), dim=1).reshape(-1).float().contiguous()
                c = self._fused = (key, w, b, out.bias.float().contiguous(), src)
        return c[1:4]

    # norm3 folded into the GEGLU GEMM (apx_linear_lnfold): one kernel less per block; APX_LN_FOLD=0 keeps the separate LayerNorm
    fold_norm = os.environ.get("APX_LN_FOLD", "1") != "0"

    def _packed_ln(self, norm: nn.LayerNorm):
        """(Wg interleaved, c1, c2, fp32 out bias) for LayerNorm(norm) folded into the GEGLU projection."""
        proj, out = self.net[0].proj, self.net[2]
        src = (proj.weight, proj.bias, out.bias, norm.weight, norm.bias)

        def build():
            inner = proj.weight.shape[0] // 2
            wg, c1, c2 = _lib.fold_layernorm(proj.weight.detach(), proj.bias.detach(), norm.weight.detach(), norm.bias.detach())
            il = lambda t: torch.stack((t[:inner], t[inner:]), dim=1).reshape((2 * inner,) + t.shape[1:]).contiguous()  # noqa: E731
            return il(wg), il(c1), il(c2), out.bias.detach().float().contiguous()

        return _cached(self, "ln_fold", src, build)

    def fused_ln(self, h, norm: nn.LayerNorm):
        """ff(norm(h)) + h on (..., dim) 16-bit CUDA tensors, the LayerNorm folded into the first GEMM."""
        wg, c1, c2, b2 = self._packed_ln(norm)
        shape = h.shape
        h2 = h.reshape(-1, shape[-1])
        h2 = h2 if h2.stride(1) == 1 else h2.contiguous()
        hidden = _lib.linear_lnfold(h2, wg, c1, c2, norm.eps, geglu=True)
        return _lib.linear(hidden, self.net[2].weight, b2, residual=h2).reshape(shape)

    def fused(self, x, residual):
        """ff(x) + residual on (..., dim) 16-bit CUDA tensors."""
        w1, b1, b2 = self._packed()
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        r2 = residual.reshape(-1, shape[-1])
        hidden = _lib.linear(x2, w1, b1, geglu=True)
        return _lib.linear(hidden, self.net[2].weight, b2, residual=r2 if r2.stride(1) == 1 else r2.contiguous()).reshape(shape)


class BasicTransformerBlock(nn.Module):
    """norm1 -> attn1 -> +res; norm2 -> attn2(ctx) -> +res; norm3 -> GEGLU FF -> +res."""

    # class-wide switch (APX_FOLD_RESIDUAL): "ip" (default) only the adapter sites, which add in place; "1" also the plain sites
    # (to_out + residual through libapx's GEMM — measured 0.37 ms per step slower than cuBLAS + add on B200 with the round-1 GEMM, 0.09 ms
    # slower with gemm16 and apx_add_layernorm (8.39 vs 8.30 ms, profiles/r3o_bench_b1_fold_residual_all.json): kept opt-in); "0" none
    fold_residual = {"0": False, "1": True}.get(os.environ.get("APX_FOLD_RESIDUAL", "ip"), "ip")
    # inference on CUDA: where norm1 / norm2 are not folded into the next projection (the sites below 1000 rows), attn1's residual add and
    # norm2 run as ONE libapx kernel (apx_add_layernorm) and norm1 as its LayerNorm-only form; APX_ADD_LN=0 keeps the ATen kernels
    add_ln = os.environ.get("APX_ADD_LN", "1") != "0"

    def __init__(self, dim: int, heads: int, dim_head: int, cross_attention_dim: Optional[int],
                 double_self_attention: bool = False, only_cross_attention: bool = False):
        super().__init__()
        self.only_cross_attention = only_cross_attention
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, cross_attention_dim if only_cross_attention else None, heads, dim_head)
        if cross_attention_dim is not None or double_self_attention:
            self.norm2 = nn.LayerNorm(dim)
            self.attn2 = Attention(dim, None if double_self_attention else cross_attention_dim, heads, dim_head)
        else:
            self.norm2 = None
            self.attn2 = None
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, h, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                cross_attention_kwargs=None):
        kw = cross_attention_kwargs or {}
        # Inference on CUDA: processors that take `residual_inout` fold this block's residual add into their last kernel (the
        # adapter sites add into h in place — h is this block's own tensor by then —, the plain sites return a new tensor).
        fold = h.is_cuda and not torch.is_grad_enabled() and not h.requires_grad and self.fold_residual
        ehs1 = encoder_hidden_states if self.only_cross_attention else None
        p1 = self.attn1.processor
        kw1 = dict(kw)
        infer = h.is_cuda and not torch.is_grad_enabled() and not h.requires_grad
        if hasattr(p1, "can_fold_norm") and p1.can_fold_norm(self.attn1, h, self.norm1):
            x, kw1["fold_norm"] = h, self.norm1          # norm1 runs inside the site's QKV projection
        elif self.add_ln and infer and h.is_contiguous() and _lib.add_layernorm_ok(h, self.norm1):
            x = _lib.add_layernorm(h, None, self.norm1)[1]   # libapx LayerNorm (one warp per row, every load in flight at once)
        else:
            x = self.norm1(h)
        a1 = None
        if fold is True and getattr(p1, "accepts_residual", False):
            h = self.attn1(x, encoder_hidden_states=ehs1, attention_mask=attention_mask, residual_inout=h, **kw1)
        else:
            a1 = self.attn1(x, encoder_hidden_states=ehs1, attention_mask=attention_mask, **kw1)   # residual add below
        if self.attn2 is None:
            if a1 is not None:
                h = a1 + h
        else:
            # the adapter hot path: attn2's input is norm2(h), its output is residual-added right here
            p2 = self.attn2.processor
            kw2 = dict(kw)
            fold2 = hasattr(p2, "can_fold_norm") and p2.can_fold_norm(self.attn2, h, self.norm2, encoder_hidden_states)
            x = None
            if a1 is not None:
                if (not fold2 and self.add_ln and infer and a1.shape == h.shape and h.is_contiguous() and _lib.add_layernorm_ok(a1, self.norm2)):
                    h, x = _lib.add_layernorm(a1, h, self.norm2)   # attn1's residual add + norm2 in one kernel (apx_add_layernorm)
                else:
                    h = a1 + h
            if fold2:
                x, kw2["fold_norm"] = h, self.norm2      # norm2 folded into the adapter site kernel / a plain site's Q(KV) projection
            elif x is None:
                if self.add_ln and infer and h.is_contiguous() and _lib.add_layernorm_ok(h, self.norm2):
                    x = _lib.add_layernorm(h, None, self.norm2)[1]   # (attn1 folded its own residual add: LayerNorm-only form)
                else:
                    x = self.norm2(h)
            if fold and getattr(p2, "accepts_residual", False) and (fold is True or getattr(p2, "folds_in_place", False)):
                h = self.attn2(x, encoder_hidden_states=encoder_hidden_states, attention_mask=encoder_attention_mask,
                               residual_inout=h.contiguous(), **kw2)
            else:
                h = self.attn2(x, encoder_hidden_states=encoder_hidden_states, attention_mask=encoder_attention_mask, **kw2) + h
        if self.ff.fold_norm and self.ff.fused_ok(h) and self.norm3.elementwise_affine:
            return self.ff.fused_ln(h, self.norm3)
        x = self.norm3(h)
        if self.ff.fused_ok(x):
            return self.ff.fused(x, h)
        return self.ff(x) + h


class Transformer2DModel(nn.Module):
    def __init__(self, heads: int, dim_head: int, in_channels: int, num_layers: int = 1,
                 cross_attention_dim: Optional[int] = None, norm_num_groups: int = 32, use_linear_projection: bool = False,
                 only_cross_attention: bool = False, double_self_attention: bool = False):
        super().__init__()
        inner = heads * dim_head
        self.use_linear_projection = use_linear_projection
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        if use_linear_projection:
            self.proj_in = nn.Linear(in_channels, inner)
            self.proj_out = nn.Linear(inner, in_channels)
        else:
            self.proj_in = nn.Conv2d(in_channels, inner, kernel_size=1)
            self.proj_out = nn.Conv2d(inner, in_channels, kernel_size=1)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim, double_self_attention, only_cross_attention)
            for _ in range(num_layers)])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, encoder_attention_mask=None,
                cross_attention_kwargs=None, return_dict: bool = False):
        if encoder_attention_mask is not None and encoder_attention_mask.ndim == 2:
            encoder_attention_mask = ((1 - encoder_attention_mask.to(hidden_states.dtype)) * -10000.0).unsqueeze(1)
        b, _, hh, ww = hidden_states.shape
        if NHWC.active and not self.use_linear_projection and NHWC.usable(hidden_states):
            return (self._forward_nhwc(hidden_states, encoder_hidden_states, attention_mask, encoder_attention_mask,
                                       cross_attention_kwargs),)
        residual = hidden_states
        h = self.norm(hidden_states)
        if not self.use_linear_projection:
            h = self.proj_in(h)
            h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, h.shape[1])
        else:
            h = self.proj_in(h.permute(0, 2, 3, 1).reshape(b, hh * ww, h.shape[1]))
        for blk in self.transformer_blocks:
            h = blk(h, attention_mask=attention_mask, encoder_hidden_states=encoder_hidden_states,
                    encoder_attention_mask=encoder_attention_mask, cross_attention_kwargs=cross_attention_kwargs)
        if not self.use_linear_projection:
            h = self.proj_out(h.reshape(b, hh, ww, -1).permute(0, 3, 1, 2).contiguous())
        else:
            h = self.proj_out(h).reshape(b, hh, ww, -1).permute(0, 3, 1, 2).contiguous()
        return (h + residual,)


    def _forward_nhwc(self, x, encoder_hidden_states, attention_mask, encoder_attention_mask, cross_attention_kwargs):
        """Channels-last route: GroupNorm -> free token view -> proj_in as a linear -> blocks -> proj_out as a linear + residual
        -> channels-last view.  No permute copies, no NCHW<->NHWC transposes, no separate conv-bias adds."""
        x = _as_cl(x)
        b, c, hh, ww = x.shape
        inner = self.proj_in.out_channels
        tok = _gn_nhwc(x, self.norm, silu=False).permute(0, 2, 3, 1).reshape(b, hh * ww, c)      # a view of the NHWC buffer
        h = F.linear(tok, self.proj_in.weight.view(inner, c), self.proj_in.bias)
        for blk in self.transformer_blocks:
            h = blk(h, attention_mask=attention_mask, encoder_hidden_states=encoder_hidden_states,
                    encoder_attention_mask=encoder_attention_mask, cross_attention_kwargs=cross_attention_kwargs)
        out = F.linear(h, self.proj_out.weight.view(c, inner), self.proj_out.bias)
        out = out + x.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
        return out.view(b, hh, ww, c).permute(0, 3, 1, 2)                                          # channels-last strides


class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, groups: int = 32, eps: float = 1e-5,
                 output_scale_factor: float = 1.0):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None
        self.output_scale_factor = output_scale_factor

    def forward(self, x, temb):
        if NHWC.active and NHWC.usable(x):
            return self._forward_nhwc(x, temb)
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(self.dropout(F.silu(self.norm2(h))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return (x + h) / self.output_scale_factor


    def _forward_nhwc(self, x, temb):
        """Channels-last route, 5 (6 with a shortcut) kernels instead of 15: conv1's bias and the time-embedding add live in
        norm2's kernel, conv2's (and the shortcut's) bias in the output add."""
        x = _as_cl(x)
        t = self.__dict__.get("_temb_proj")          # this forward's slice of the UNet's batched time-embedding GEMM
        if t is None:
            t = self.time_emb_proj(F.silu(temb))
        h = F.conv2d(_gn_nhwc(x, self.norm1, silu=True), _cl_weight(self.conv1), None, padding=self.conv1.padding)
        h = _gn_nhwc(h, self.norm2, silu=True, chan_bias=_f32(self.conv1, "bias"), bc_add=t)
        h = F.conv2d(h, _cl_weight(self.conv2), None, padding=self.conv2.padding)
        if self.conv_shortcut is not None:
            x = F.conv2d(x, _cl_weight(self.conv_shortcut), None)
            bias = _cached(self, "bias2", (self.conv2.bias, self.conv_shortcut.bias),
                           lambda: (self.conv2.bias.detach().float() + self.conv_shortcut.bias.detach().float()).contiguous())
        else:
            bias = _f32(self.conv2, "bias")
        if x.shape[1] % 8:
            return (x + h + bias.to(h.dtype)[None, :, None, None]) / self.output_scale_factor
        return _lib.add_bias_nhwc(_as_cl(x), _as_cl(h), bias, 1.0 / self.output_scale_factor)


def _conv_nhwc(conv: nn.Conv2d, x):
    return F.conv2d(_as_cl(x), _cl_weight(conv), conv.bias, stride=conv.stride, padding=conv.padding)


class Downsample2D(nn.Module):
    def __init__(self, channels: int, padding: int = 1):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=padding)

    def forward(self, x):
        if NHWC.active and NHWC.usable(x):
            return _conv_nhwc(self.conv, x)
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x, output_size=None):
        if output_size is None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        else:
            x = F.interpolate(x, size=tuple(output_size), mode="nearest")
        if NHWC.active and NHWC.usable(x):
            return _conv_nhwc(self.conv, x)
        return self.conv(x)


class Timesteps(nn.Module):
    def __init__(self, num_channels: int, flip_sin_to_cos: bool = True, downscale_freq_shift: float = 0.0):
        super().__init__()
        self.num_channels, self.flip, self.shift = num_channels, flip_sin_to_cos, downscale_freq_shift

    def forward(self, timesteps):
        half = self.num_channels // 2
        exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / (half - self.shift)
        emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
        sin, cos = torch.sin(emb), torch.cos(emb)
        return torch.cat([cos, sin], dim=-1) if self.flip else torch.cat([sin, cos], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample, condition=None):
        return self.linear_2(F.silu(self.linear_1(sample)))


# ---------------------------------------------------------------------------------------------------------
# reference wiring (pipeline/modeling_audioldm2.py)
# ---------------------------------------------------------------------------------------------------------
def _route(cross_dims, idx, ehs, mask, ehs1, mask1):
    """Per-slot context routing of the reference blocks (modeling_audioldm2.py:1141-1149): slots 0-1 with a cross dim
    read ``encoder_hidden_states`` (768-d GPT-2 + AudioMAE tokens), slots 2-3 ``encoder_hidden_states_1`` (T5),
    a ``None`` slot is (double) self-attention."""
    if cross_dims[idx] is None:
        return None, None
    return (ehs, mask) if idx <= 1 else (ehs1, mask1)


class _CrossAttnBlockBase(nn.Module):
    has_cross_attention = True

    def _make_attentions(self, n_layers, channels, heads, cross_dims, tlayers, groups, linear, only_cross=False):
        return nn.ModuleList([
            Transformer2DModel(heads, channels // heads, channels, tlayers, cd, groups, linear, only_cross,
                               double_self_attention=cd is None)
            for _ in range(n_layers) for cd in cross_dims])

    def _run_attentions(self, i, h, ehs, mask, ehs1, mask1, attention_mask, kw):
        n = len(self.cross_attention_dim)
        for idx in range(n):
            e, m = _route(self.cross_attention_dim, idx, ehs, mask, ehs1, mask1)
            h = self.attentions[i * n + idx](h, encoder_hidden_states=e, attention_mask=attention_mask,
                                             encoder_attention_mask=m, cross_attention_kwargs=kw)[0]
        return h


class CrossAttnDownBlock2D(_CrossAttnBlockBase):
    """modeling_audioldm2.py:1003-1173."""

    def __init__(self, in_channels, out_channels, temb_channels, num_layers, transformer_layers_per_block, heads,
                 cross_attention_dim, groups, eps, add_downsample, downsample_padding, use_linear_projection):
        super().__init__()
        self.cross_attention_dim = tuple(cross_attention_dim)
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels,
                                                    groups, eps) for i in range(num_layers)])
        self.attentions = self._make_attentions(num_layers, out_channels, heads, self.cross_attention_dim,
                                                transformer_layers_per_block, groups, use_linear_projection)
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels, downsample_padding)]) if add_downsample else None

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs=None, encoder_attention_mask=None, encoder_hidden_states_1=None,
                encoder_attention_mask_1=None):
        if encoder_hidden_states_1 is None:
            encoder_hidden_states_1, encoder_attention_mask_1 = encoder_hidden_states, encoder_attention_mask
        out = ()
        for i, resnet in enumerate(self.resnets):
            hidden_states = resnet(hidden_states, temb)
            hidden_states = self._run_attentions(i, hidden_states, encoder_hidden_states, encoder_attention_mask,
                                                 encoder_hidden_states_1, encoder_attention_mask_1, attention_mask,
                                                 cross_attention_kwargs)
            out += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            out += (hidden_states,)
        return hidden_states, out


class DownBlock2D(nn.Module):
    has_cross_attention = False

    def __init__(self, in_channels, out_channels, temb_channels, num_layers, groups, eps, add_downsample,
                 downsample_padding):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels,
                                                    groups, eps) for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels, downsample_padding)]) if add_downsample else None

    def forward(self, hidden_states, temb=None):
        out = ()
        for resnet in self.resnets:
            hidden_states = resnet(hidden_states, temb)
            out += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            out += (hidden_states,)
        return hidden_states, out


class UNetMidBlock2DCrossAttn(_CrossAttnBlockBase):
    """modeling_audioldm2.py:1176-1335: resnet, then per layer (4 transformers, resnet)."""

    def __init__(self, in_channels, temb_channels, num_layers, transformer_layers_per_block, heads, cross_attention_dim,
                 groups, eps, output_scale_factor, use_linear_projection):
        super().__init__()
        self.cross_attention_dim = tuple(cross_attention_dim)
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels, in_channels, temb_channels, groups, eps,
                                                    output_scale_factor) for _ in range(num_layers + 1)])
        self.attentions = self._make_attentions(num_layers, in_channels, heads, self.cross_attention_dim,
                                                transformer_layers_per_block, groups, use_linear_projection)

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs=None, encoder_attention_mask=None, encoder_hidden_states_1=None,
                encoder_attention_mask_1=None):
        if encoder_hidden_states_1 is None:
            encoder_hidden_states_1, encoder_attention_mask_1 = encoder_hidden_states, encoder_attention_mask
        hidden_states = self.resnets[0](hidden_states, temb)
        for i in range(len(self.resnets) - 1):
            hidden_states = self._run_attentions(i, hidden_states, encoder_hidden_states, encoder_attention_mask,
                                                 encoder_hidden_states_1, encoder_attention_mask_1, attention_mask,
                                                 cross_attention_kwargs)
            hidden_states = self.resnets[i + 1](hidden_states, temb)
        return hidden_states


class CrossAttnUpBlock2D(_CrossAttnBlockBase):
    """modeling_audioldm2.py:1338-1513."""

    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels, num_layers,
                 transformer_layers_per_block, heads, cross_attention_dim, groups, eps, add_upsample,
                 use_linear_projection):
        super().__init__()
        self.cross_attention_dim = tuple(cross_attention_dim)
        resnets = []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            cin = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(cin + skip, out_channels, temb_channels, groups, eps))
        self.resnets = nn.ModuleList(resnets)
        self.attentions = self._make_attentions(num_layers, out_channels, heads, self.cross_attention_dim,
                                                transformer_layers_per_block, groups, use_linear_projection)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None,
                cross_attention_kwargs=None, upsample_size=None, attention_mask=None, encoder_attention_mask=None,
                encoder_hidden_states_1=None, encoder_attention_mask_1=None):
        if encoder_hidden_states_1 is None:
            encoder_hidden_states_1, encoder_attention_mask_1 = encoder_hidden_states, encoder_attention_mask
        for i, resnet in enumerate(self.resnets):
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = resnet(torch.cat([hidden_states, res], dim=1), temb)
            hidden_states = self._run_attentions(i, hidden_states, encoder_hidden_states, encoder_attention_mask,
                                                 encoder_hidden_states_1, encoder_attention_mask_1, attention_mask,
                                                 cross_attention_kwargs)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


class UpBlock2D(nn.Module):
    has_cross_attention = False

    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels, num_layers, groups, eps,
                 add_upsample):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            cin = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(cin + skip, out_channels, temb_channels, groups, eps))
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, upsample_size=None):
        for resnet in self.resnets:
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = resnet(torch.cat([hidden_states, res], dim=1), temb)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


AUDIOLDM2_LARGE_CONFIG = dict(
    in_channels=8, out_channels=8,
    down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
    up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
    block_out_channels=(128, 256, 384, 640), layers_per_block=2, transformer_layers_per_block=2,
    cross_attention_dim=(None, 768, 1024, None), attention_head_dim=8, norm_num_groups=32, norm_eps=1e-5,
    downsample_padding=1, mid_block_scale_factor=1.0, use_linear_projection=False, flip_sin_to_cos=True, freq_shift=0,
)


class AudioLDM2UNet2DConditionModel(nn.Module):
    """``AudioLDM2UNet2DConditionModel`` (modeling_audioldm2.py:150-873) with the geometry of ``cvssp/audioldm2-large``
    as inferred from the shipped adapter weights.  ``cross_attention_dim`` is the per-slot tuple of ONE block
    (every attention-bearing block uses the same four slots)."""

    def __init__(self, **overrides):
        super().__init__()
        cfg = dict(AUDIOLDM2_LARGE_CONFIG)
        cfg.update(overrides)
        self.config = SimpleNamespace(**cfg)
        c = self.config
        boc = tuple(c.block_out_channels)
        heads = c.attention_head_dim  # diffusers' naming quirk: attention_head_dim is the NUMBER of heads (:273-279)
        cross = tuple(c.cross_attention_dim)
        temb = boc[0] * 4
        self.conv_in = nn.Conv2d(c.in_channels, boc[0], 3, padding=1)
        self.time_proj = Timesteps(boc[0], c.flip_sin_to_cos, c.freq_shift)
        self.time_embedding = TimestepEmbedding(boc[0], temb)

        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, kind in enumerate(c.down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            final = i == len(boc) - 1
            if kind == "CrossAttnDownBlock2D":
                blk = CrossAttnDownBlock2D(in_ch, out_ch, temb, c.layers_per_block, c.transformer_layers_per_block, heads,
                                           cross, c.norm_num_groups, c.norm_eps, not final, c.downsample_padding,
                                           c.use_linear_projection)
            elif kind == "DownBlock2D":
                blk = DownBlock2D(in_ch, out_ch, temb, c.layers_per_block, c.norm_num_groups, c.norm_eps, not final,
                                  c.downsample_padding)
            else:
                raise ValueError(f"{kind} does not exist.")
            self.down_blocks.append(blk)

        self.mid_block = UNetMidBlock2DCrossAttn(boc[-1], temb, 1, c.transformer_layers_per_block, heads, cross,
                                                 c.norm_num_groups, c.norm_eps, c.mid_block_scale_factor,
                                                 c.use_linear_projection)

        self.up_blocks = nn.ModuleList()
        self.num_upsamplers = 0
        rev = list(reversed(boc))
        out_ch = rev[0]
        for i, kind in enumerate(c.up_block_types):
            final = i == len(boc) - 1
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            if not final:
                self.num_upsamplers += 1
            if kind == "CrossAttnUpBlock2D":
                blk = CrossAttnUpBlock2D(in_ch, out_ch, prev, temb, c.layers_per_block + 1, c.transformer_layers_per_block,
                                         heads, cross, c.norm_num_groups, c.norm_eps, not final, c.use_linear_projection)
            elif kind == "UpBlock2D":
                blk = UpBlock2D(in_ch, out_ch, prev, temb, c.layers_per_block + 1, c.norm_num_groups, c.norm_eps, not final)
            else:
                raise ValueError(f"{kind} does not exist.")
            self.up_blocks.append(blk)

        self.conv_norm_out = nn.GroupNorm(c.norm_num_groups, boc[0], eps=c.norm_eps)
        self.conv_out = nn.Conv2d(boc[0], c.out_channels, 3, padding=1)

    # ---- processor registry (modeling_audioldm2.py:517-573) ------------------------------------------------
    @property
    def attn_processors(self) -> Dict[str, nn.Module]:
        procs: Dict[str, nn.Module] = {}

        def walk(name, module):
            if hasattr(module, "set_processor"):
                procs[f"{name}.processor"] = module.processor
            for sub, child in module.named_children():
                if sub == "processor":
                    continue
                walk(f"{name}.{sub}", child)

        for name, module in self.named_children():
            walk(name, module)
        return procs

    def set_attn_processor(self, processor) -> None:
        count = len(self.attn_processors.keys())
        if isinstance(processor, dict) and len(processor) != count:
            raise ValueError(f"A dict of processors was passed, but the number of processors {len(processor)} does not match "
                             f"the number of attention layers: {count}. Please make sure to pass {count} processor classes.")

        def walk(name, module):
            if hasattr(module, "set_processor"):
                module.set_processor(processor if not isinstance(processor, dict) else processor.pop(f"{name}.processor"))
            for sub, child in module.named_children():
                if sub == "processor":
                    continue
                walk(f"{name}.{sub}", child)

        for name, module in self.named_children():
            walk(name, module)

    def invalidate_caches(self) -> None:
        """Drop every derived tensor the inference path keeps next to the weights (fp32 / channels-last copies, concatenated and
        LayerNorm-folded projections, packed GEGLU weights, the processors' context K/V and packed-weight blobs, additive masks).
        The caches are keyed on tensor identity + ``_version``; call this after edits the keys cannot see — ``param.data`` writes
        (8-bit / fused optimisers, EMA copies), raw kernels or graph replays that rewrite a keyed tensor in place."""
        for m in self.modules():
            m.__dict__.pop(_CACHE_ATTR, None)
            if isinstance(m, FeedForward):
                m._fused = None
        for proc in self.attn_processors.values():
            if hasattr(proc, "invalidate_cache"):
                proc.invalidate_cache()

    # ---- forward (modeling_audioldm2.py:663-873) --------------------------------------------------------------
    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, timestep_cond=None,
                attention_mask=None, cross_attention_kwargs=None, encoder_attention_mask=None, return_dict: bool = False,
                encoder_hidden_states_1=None, encoder_attention_mask_1=None):
        forward_upsample_size = any(s % (2 ** self.num_upsamplers) != 0 for s in sample.shape[-2:])
        encoder_attention_mask = self._additive_mask(encoder_attention_mask, sample.dtype, "m0")
        encoder_attention_mask_1 = self._additive_mask(encoder_attention_mask_1, sample.dtype, "m1")

        timesteps = timestep
        if not torch.is_tensor(timesteps):
            dt = torch.float64 if isinstance(timestep, float) else torch.int64
            timesteps = torch.tensor([timesteps], dtype=dt, device=sample.device)
        elif timesteps.dim() == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0])
        emb = self.time_embedding(self.time_proj(timesteps).to(dtype=sample.dtype), timestep_cond)

        if NHWC.usable(sample) and not NHWC.active:
            NHWC.active = True
            try:
                return self._forward_trunk_nhwc(sample, emb, encoder_hidden_states, attention_mask, cross_attention_kwargs,
                                                encoder_attention_mask, encoder_hidden_states_1, encoder_attention_mask_1,
                                                forward_upsample_size, return_dict)
            finally:
                NHWC.active = False
                for r in self._resnets():
                    r.__dict__.pop("_temb_proj", None)
        return self._forward_trunk(sample, emb, encoder_hidden_states, attention_mask, cross_attention_kwargs, encoder_attention_mask,
                                   encoder_hidden_states_1, encoder_attention_mask_1, forward_upsample_size, return_dict)

    def _additive_mask(self, mask, dtype, tag):
        """(B, L) keep-mask -> additive (B, 1, L) bias (modeling_audioldm2.py:741-747).  Under inference the result is cached per mask
        identity + version, so every step hands the processors the SAME tensor and their per-site mask caches hit."""
        if mask is None:
            return None
        if torch.is_grad_enabled() or not AttnProcessor2_0.cache_context:
            return ((1 - mask.to(dtype)) * -10000.0).unsqueeze(1)
        cache = self.__dict__.setdefault(_CACHE_ATTR, {})
        key = (mask.data_ptr(), mask._version, tuple(mask.shape), mask.dtype, dtype)
        hit = cache.get(tag)
        if hit is None or hit[0] != key or hit[2] is not mask:
            hit = cache[tag] = (key, ((1 - mask.to(dtype)) * -10000.0).unsqueeze(1), mask)
        return hit[1]

    def _resnets(self):
        r = self.__dict__.get("_resnet_list")
        if r is None:
            r = self.__dict__["_resnet_list"] = [m for m in self.modules() if isinstance(m, ResnetBlock2D)]
        return r

    def _forward_trunk_nhwc(self, sample, emb, encoder_hidden_states, attention_mask, cross_attention_kwargs, encoder_attention_mask,
                            encoder_hidden_states_1, encoder_attention_mask_1, forward_upsample_size, return_dict):
        """The same trunk in channels-last memory (see NHWC).  All time-embedding projections of the forward are one GEMM over the
        concatenated ``time_emb_proj`` weights; each ResnetBlock2D picks its column slice."""
        resnets = self._resnets()
        params = [p for r in resnets for p in (r.time_emb_proj.weight, r.time_emb_proj.bias)]
        w_all, b_all = _cached(self, "temb_all", params,
                               lambda: (torch.cat([r.time_emb_proj.weight.detach() for r in resnets], 0).contiguous(),
                                        torch.cat([r.time_emb_proj.bias.detach() for r in resnets], 0).contiguous()))
        t_all = F.linear(F.silu(emb), w_all, b_all)
        off = 0
        for r in resnets:
            n = r.time_emb_proj.out_features
            r.__dict__["_temb_proj"] = t_all[:, off:off + n]
            off += n
        sample = _as_cl(sample)
        out = self._forward_trunk(sample, emb, encoder_hidden_states, attention_mask, cross_attention_kwargs, encoder_attention_mask,
                                  encoder_hidden_states_1, encoder_attention_mask_1, forward_upsample_size, return_dict=False)[0]
        out = out.contiguous()          # the caller's layout (NCHW), 8 channels
        return (out,) if not return_dict else SimpleNamespace(sample=out)

    def _forward_trunk(self, sample, emb, encoder_hidden_states, attention_mask, cross_attention_kwargs, encoder_attention_mask,
                       encoder_hidden_states_1, encoder_attention_mask_1, forward_upsample_size, return_dict):
        upsample_size = None
        nhwc = NHWC.active and NHWC.usable(sample)
        sample = _conv_nhwc(self.conv_in, sample) if nhwc else self.conv_in(sample)
        skips: Tuple[torch.Tensor, ...] = (sample,)
        kw = dict(encoder_hidden_states=encoder_hidden_states, attention_mask=attention_mask,
                  cross_attention_kwargs=cross_attention_kwargs, encoder_attention_mask=encoder_attention_mask,
                  encoder_hidden_states_1=encoder_hidden_states_1, encoder_attention_mask_1=encoder_attention_mask_1)
        for blk in self.down_blocks:
            if blk.has_cross_attention:
                sample, res = blk(hidden_states=sample, temb=emb, **kw)
            else:
                sample, res = blk(hidden_states=sample, temb=emb)
            skips += res
        sample = self.mid_block(sample, emb, **kw)
        for i, blk in enumerate(self.up_blocks):
            final = i == len(self.up_blocks) - 1
            n = len(blk.resnets)
            res, skips = skips[-n:], skips[:-n]
            if not final and forward_upsample_size:
                upsample_size = skips[-1].shape[2:]
            if blk.has_cross_attention:
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res, upsample_size=upsample_size, **kw)
            else:
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res, upsample_size=upsample_size)
        if nhwc:
            sample = _conv_nhwc(self.conv_out, _gn_nhwc(sample, self.conv_norm_out, silu=True))
        else:
            sample = self.conv_out(F.silu(self.conv_norm_out(sample)))
        if not return_dict:
            return (sample,)
        return SimpleNamespace(sample=sample)


# ---------------------------------------------------------------------------------------------------------
# processor injection, as the reference's entry points do it
# ---------------------------------------------------------------------------------------------------------
SLOT_CROSS_DIMS = [None, None, 768, 768, 1024, 1024, None, None]  # inference.py:16


def build_attn_processors(unet, ip_cls, plain_cls=AttnProcessor2_0, ap_scale: float = 1.0, num_tokens: int = 8,
                          do_copy: bool = False, device=None, dtype=None) -> Dict[str, nn.Module]:
    """The processor dict of inference.py:15-49 / train_apadapter_v2.py:613-648: ``attn1`` -> plain; ``attn2`` slots cycle
    through ``[None, None, 768, 768, 1024, 1024, None, None]`` and only the 768-d ones get ``ip_cls``."""
    procs: Dict[str, nn.Module] = {}
    layer_num = 0
    boc = list(unet.config.block_out_channels)
    for name in unet.attn_processors.keys():
        if name.startswith("mid_block"):
            hidden = boc[-1]
        elif name.startswith("up_blocks"):
            hidden = list(reversed(boc))[int(name[len("up_blocks.")])]
        else:
            hidden = boc[int(name[len("down_blocks.")])]
        if name.endswith("attn1.processor"):
            procs[name] = plain_cls()
            continue
        cd = SLOT_CROSS_DIMS[layer_num % 8]
        layer_num += 1
        if cd == 768:
            p = ip_cls(hidden_size=hidden, name=name, cross_attention_dim=cd, scale=ap_scale, num_tokens=num_tokens,
                       do_copy=do_copy)
            procs[name] = p.to(device=device, dtype=dtype) if (device is not None or dtype is not None) else p
        else:
            procs[name] = plain_cls()
    return procs
