"""Drop-in attention processors backed by the sm_100a kernels.

Mirrors the reference interface (``APadapter/ap_adapter/attention_processor.py``): same class names, constructor
arguments, attribute names (``to_k_ip``, ``to_v_ip``, ``scale``, ``num_tokens``, ``name`` ...), ``__call__``
signature and checkpoint formats, so ``inference.py:22-59`` / ``train_apadapter_v2.py:613-669`` can build and
register them unchanged with ``unet.set_attn_processor``.

Differences that are deliberate (SURVEY.md §8 a-2'):
* ``scale != 1.0`` passed to ``__call__`` warns instead of hitting the reference's undefined ``logger`` (NameError,
  attention_processor.py:356-357).
* a non-None ``attention_mask`` on the adapter site raises: the pipeline never passes one for the 768-d context
  (pipeline_audioldm2.py:1010-1017) and the reference's ``split`` hack (:424-428) is not reproduced.
* the context K/V projections (attention_processor.py:406-407,435-436) are timestep-invariant; they are computed
  once per (context, weights) and cached on the processor instead of being recomputed every denoise step.
* with gradients enabled (adapter training, train_apadapter_v2.py:892-979) the call goes through ``_IPAttnSiteFn``:
  same forward kernels, ``apx_ipattn_bwd`` for dX / d to_k_ip.weight / d to_v_ip.weight; every other weight is frozen.
"""
from __future__ import annotations

import os
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


def _copied_weight_dir() -> str:
    """``../copied_cross_attention`` relative to the package, as the reference resolves it
    (attention_processor.py:330-335); overridable for deployments that keep the files elsewhere."""
    return os.environ.get(
        "APX_COPIED_CROSS_ATTENTION",
        os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "copied_cross_attention"))


class AttnProcessor2_0(nn.Module):
    """Plain scaled-dot-product attention processor for the non-adapter sites (self-attention, T5 cross-attention).

    Reference: attention_processor.py:199-294.  These 224 sites are outside the adapter hot path
    and stay a thin PyTorch module: cuBLAS projections + torch SDPA, with the reference's mask handling.
    """

    # inference on CUDA: to_q/to_k/to_v of a self-attention site (to_k/to_v of a cross-attention site) run as ONE projection over
    # the concatenated weight (what diffusers' fuse_qkv_projections() does); APX_QKV_FUSED=0 keeps the three separate GEMMs
    fuse_projections = os.environ.get("APX_QKV_FUSED", "1") != "0"
    # inference on CUDA: prepared attention masks (and, with fuse_projections, the K / V of a cross-attention context) are cached per
    # tensor identity + version instead of being rebuilt at every site and step; False = rebuild every call, as the reference does
    cache_context = True
    # inference on CUDA: the LayerNorm in front of the site (norm1 / norm2 of the transformer block) is folded into the Q(KV)
    # projection (apx_linear_lnfold) when the block hands over the un-normalised stream (`fold_norm=`); only for sites with at
    # least this many query rows (measured on B200 at B = 1, ms per step: fold everywhere 10.34, >= 400 rows 10.44, >= 1000 rows 9.94,
    # nowhere 10.11 — below ~1000 rows the unsplit tcgen05 GEMM loses to cuBLAS's split-K + a separate LayerNorm); APX_LN_FOLD_QKV=0 disables
    fold_norm_min_rows = int(os.environ.get("APX_LN_FOLD_QKV", "1000"))
    # inference on CUDA: the unmasked SDPA of a site with at least this many keys runs in libapx (apx_flash_attention) instead of
    # torch's cuDNN SDPA; APX_FLASH_ATTN=0 disables, APX_FLASH_ATTN=n sets the threshold
    flash_min_keys = (lambda v: 1 << 30 if v == "0" else int(v))(os.environ.get("APX_FLASH_ATTN", "200"))
    # ... and only while its grid (one CTA per 128-query tile, head and batch element) fits one wave of the 148 SMs: measured on B200,
    # 11.3 vs 14.5 us (cuDNN) at B = 2, S = 1000, C = 256, but 76 vs 49 us at B = 16 (one CTA per SM with 232-register softmax
    # threads: once there is more than one wave cuDNN's kernel wins)
    flash_max_ctas = int(os.environ.get("APX_FLASH_MAX_CTAS", "160"))
    # ... and for the head sizes whose rows are exactly one TMA swizzle span (64 / 128 bytes: whole-row Q / K / V loads; at d = 48 the
    # 16-byte-piece tile loads leave it at 7.3 us against cuDNN's 6.9 at S = 252)
    flash_head_dims = (32, 64)

    def __init__(self, hidden_size=None, cross_attention_dim=None):
        super().__init__()
        if not hasattr(F, "scaled_dot_product_attention"):
            raise ImportError("AttnProcessor2_0 requires PyTorch 2.0, to use it, please upgrade PyTorch to 2.0.")
        self._cat = None

    def invalidate_cache(self):
        """Drop the concatenated / LayerNorm-folded projection weights, the cached context K / V, the prepared mask and the fp32
        bias copy (all keyed on tensor identity + ``_version``; needed after edits that do not bump the version)."""
        self._cat = None
        for name in ("_fold", "_kv", "_mk", "_bo"):
            self.__dict__.pop(name, None)

    def _cat_weight(self, mods):
        # keyed on the weight tensors themselves (kept referenced, so an address cannot be recycled under the cache) + versions
        ws = tuple(m.weight for m in mods)
        key = tuple((w._version, w.data_ptr(), w.dtype) for w in ws)
        c = self._cat
        if c is None or c[0] != key or len(c[2]) != len(ws) or any(a is not b for a, b in zip(c[2], ws)):
            c = self._cat = (key, torch.cat([w.detach() for w in ws], dim=0).contiguous(), ws)
        return c[1]

    def can_fold_norm(self, attn, hidden_states, norm, encoder_hidden_states=None) -> bool:
        """True when this call may take the un-normalised stream and fold `norm` into its first projection."""
        n = self.fold_norm_min_rows
        return (n > 0 and self.fuse_projections and hidden_states.is_cuda and not torch.is_grad_enabled() and hidden_states.dim() == 3
                and hidden_states.dtype in (torch.float16, torch.bfloat16) and attn.to_q.weight.dtype == hidden_states.dtype
                and attn.to_q.bias is None and attn.to_k.bias is None and attn.to_v.bias is None and norm.elementwise_affine
                and hidden_states.shape[-1] % 64 == 0 and attn.to_q.out_features % 64 == 0
                and hidden_states.shape[0] * hidden_states.shape[1] >= n
                and attn.spatial_norm is None and attn.group_norm is None)

    def _folded(self, mods, norm):
        """(Wg, c1, c2) of the concatenated projection `mods` with `norm` folded in, cached per weight version."""
        src = tuple(m.weight for m in mods) + (norm.weight, norm.bias)
        key = tuple((w._version, w.data_ptr(), w.dtype) for w in src)
        c = getattr(self, "_fold", None)
        if c is None or c[0] != key or len(c[2]) != len(src) or any(a is not b for a, b in zip(c[2], src)):
            w = torch.cat([m.weight.detach() for m in mods], dim=0)
            c = self._fold = (key, _lib.fold_layernorm(w, None, norm.weight.detach(), norm.bias.detach()), src)
        return c[1]

    def _project(self, attn, hidden_states, encoder_hidden_states, self_attention, fold_norm=None):
        """(query, key, value) as (B, L, C) tensors (possibly strided views of one fused projection)."""
        fused = (self.fuse_projections and hidden_states.is_cuda and not torch.is_grad_enabled()
                 and attn.to_q.bias is None and attn.to_k.bias is None and attn.to_v.bias is None)
        if fold_norm is not None:
            # hidden_states is the UN-normalised stream: LayerNorm folded into the projection GEMM (one kernel instead of two)
            b, s, c = hidden_states.shape
            mods = (attn.to_q, attn.to_k, attn.to_v) if self_attention else (attn.to_q,)
            wg, c1, c2 = self._folded(mods, fold_norm)
            out = _lib.linear_lnfold(hidden_states.reshape(b * s, c), wg, c1, c2, fold_norm.eps).view(b, s, -1)
            if self_attention:
                return out.chunk(3, dim=-1)
            hidden_states, query = None, out
        if fused and self_attention:
            return F.linear(hidden_states, self._cat_weight((attn.to_q, attn.to_k, attn.to_v))).chunk(3, dim=-1)
        if fold_norm is None:
            query = attn.to_q(hidden_states)
        if fused and encoder_hidden_states.dtype == attn.to_k.weight.dtype:
            # cross-attention context (T5 embeddings): K / V do not depend on the timestep — projected once per (context, weights)
            # identity + version and reused by every denoise step, like the adapter sites' context cache
            ws = (attn.to_k.weight, attn.to_v.weight)
            ckey = (encoder_hidden_states.data_ptr(), encoder_hidden_states._version, tuple(encoder_hidden_states.shape),
                    encoder_hidden_states.dtype) + tuple((w.data_ptr(), w._version) for w in ws)
            c = getattr(self, "_kv", None) if self.cache_context else None
            if c is None or c[0] != ckey or c[3] is not encoder_hidden_states or any(a is not b for a, b in zip(c[4], ws)):
                key, value = F.linear(encoder_hidden_states, self._cat_weight((attn.to_k, attn.to_v))).chunk(2, dim=-1)
                c = self._kv = (ckey, key, value, encoder_hidden_states, ws)
            return query, c[1], c[2]
        return query, attn.to_k(encoder_hidden_states), attn.to_v(encoder_hidden_states)   # (never reached with fold_norm: it requires `fused`)

    def _mask(self, attn, attention_mask, sequence_length, batch_size):
        """prepare_attention_mask + the per-head view (reference :245-249); cached per mask identity under inference: the T5
        mask is the same tensor at every step and every site."""
        if torch.is_grad_enabled() or not attention_mask.is_cuda or not self.cache_context:
            m = attn.prepare_attention_mask(attention_mask, sequence_length, batch_size)
            return m.view(batch_size, attn.heads, -1, m.shape[-1])
        key = (attention_mask.data_ptr(), attention_mask._version, tuple(attention_mask.shape), attention_mask.dtype, sequence_length,
               batch_size, attn.heads)
        c = getattr(self, "_mk", None)
        if c is None or c[0] != key or c[2] is not attention_mask:
            m = attn.prepare_attention_mask(attention_mask, sequence_length, batch_size)
            c = self._mk = (key, m.view(batch_size, attn.heads, -1, m.shape[-1]), attention_mask)
        return c[1]

    accepts_residual = True   # __call__ takes `residual_inout` (the restated transformer block passes its residual stream)

    def _bias_f32(self, bo):
        key = (bo._version, bo.data_ptr(), bo.dtype)
        c = getattr(self, "_bo", None)
        if c is None or c[0] != key or c[2] is not bo:
            c = self._bo = (key, bo.detach().float().contiguous(), bo)
        return c[1]

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, residual_inout=None,
                 fold_norm=None):
        """`fold_norm`: the block's LayerNorm, when `hidden_states` is the un-normalised stream (see can_fold_norm)."""
        residual = hidden_states
        if attn.spatial_norm is not None:
            hidden_states = attn.spatial_norm(hidden_states, temb)
        input_ndim = hidden_states.ndim
        if input_ndim == 4:
            batch_size, channel, height, width = hidden_states.shape
            hidden_states = hidden_states.view(batch_size, channel, height * width).transpose(1, 2)
        if encoder_hidden_states is not None and encoder_hidden_states.dim() < 3:
            encoder_hidden_states = encoder_hidden_states.unsqueeze(0)
        batch_size, sequence_length, _ = (
            hidden_states.shape if encoder_hidden_states is None else encoder_hidden_states.shape)
        if attention_mask is not None:
            attention_mask = self._mask(attn, attention_mask, sequence_length, batch_size)
        if attn.group_norm is not None:
            hidden_states = attn.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)

        self_attention = encoder_hidden_states is None
        if self_attention:
            encoder_hidden_states = hidden_states
        elif attn.norm_cross:
            encoder_hidden_states = attn.norm_encoder_hidden_states(encoder_hidden_states)
        query, key, value = self._project(attn, hidden_states, encoder_hidden_states, self_attention, fold_norm)
        head_dim = key.shape[-1] // attn.heads
        if (attention_mask is None and query.is_cuda and not torch.is_grad_enabled() and key.shape[1] >= self.flash_min_keys
                and batch_size * attn.heads * ((query.shape[1] + 127) // 128) <= self.flash_max_ctas and query.dtype in (torch.float16, torch.bfloat16) and query.dim() == 3 and key.dtype == query.dtype
                and head_dim in self.flash_head_dims and query.stride(2) == 1 and key.stride(2) == 1
                and value.stride(2) == 1):
            # the unmasked SDPA (reference :267-274) in libapx: keys streamed through tcgen05 in blocks of 128, straight from the
            # (possibly strided) q / k / v slices of the fused projection; the output is already token-major [B, S, C]
            hidden_states = _lib.flash_attention(query, key, value, attn.heads)
        else:
            query = query.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
            key = key.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
            value = value.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
            hidden_states = F.scaled_dot_product_attention(query, key, value, attn_mask=attention_mask, dropout_p=0.0,
                                                           is_causal=False)
            hidden_states = hidden_states.transpose(1, 2).reshape(batch_size, -1, attn.heads * head_dim).to(query.dtype)
        out_proj = attn.to_out[0]
        if (residual_inout is not None and input_ndim == 3 and hidden_states.is_cuda and not torch.is_grad_enabled()
                and hidden_states.dtype in (torch.float16, torch.bfloat16) and out_proj.weight.dtype == hidden_states.dtype
                and not attn.residual_connection and attn.rescale_output_factor == 1.0
                and out_proj.in_features % 64 == 0 and out_proj.out_features % 64 == 0
                and residual_inout.dtype == hidden_states.dtype and residual_inout.shape[-1] == out_proj.out_features
                and not (attn.to_out[1].training and attn.to_out[1].p > 0)):
            # to_out + bias + the block's residual add in one tcgen05 GEMM (16-bit residual epilogue); a new tensor is returned
            bias = self._bias_f32(out_proj.bias) if out_proj.bias is not None else None
            res2 = residual_inout.reshape(-1, out_proj.out_features)
            out = _lib.linear(hidden_states.reshape(-1, out_proj.in_features), out_proj.weight, bias,
                              residual=res2 if res2.stride(1) == 1 else res2.contiguous())
            return out.view(residual_inout.shape)
        hidden_states = out_proj(hidden_states)
        hidden_states = attn.to_out[1](hidden_states)
        if input_ndim == 4:
            hidden_states = hidden_states.transpose(-1, -2).reshape(batch_size, channel, height, width)
        if attn.residual_connection:
            hidden_states = hidden_states + residual
        if attn.rescale_output_factor != 1.0:   # x / 1.0 is x bit for bit; skipping it drops one kernel per site
            hidden_states = hidden_states / attn.rescale_output_factor
        return hidden_states if residual_inout is None else hidden_states + residual_inout


class _IPAttnSiteFn(torch.autograd.Function):
    """One adapter site with gradients for adapter-only training (reference: train_apadapter_v2.py:892-979 backpropagates
    through IPAttnProcessor2_0.__call__ with every UNet weight frozen).  Forward = the same libapx calls as inference
    (context cache rebuilt from the current adapter weights, then apx_ipattn_fwd); backward = apx_ipattn_bwd, which returns
    dX and the gradients of to_k_ip.weight / to_v_ip.weight."""

    @staticmethod
    def forward(ctx_, x, enc, wk_ip, wv_ip, proc, attn, heads, lt, la):
        lib = _lib.load()
        dtype = x.dtype
        # training: the adapter weights change every step, and optimisers that write through `.data` / their own kernels
        # (8-bit Adam, fused Adam, EMA copies) do not bump `_version` — always rebuild the context K/V from the live weights
        cache = proc._context_cache(attn, enc, heads, dtype, force=True)
        wq, wo, wpack = proc._site_weights(attn, heads, dtype)
        bo = attn.to_out[0].bias
        bo = proc._bias_f32(bo) if bo is not None else None
        bc, s, c = x.shape
        y = torch.empty_like(x)
        wsb = lib.apx_ipattn_workspace_bytes(bc, s, c)
        work = _lib.workspace(wsb, x.device, "ipattn")
        with torch.cuda.device(x.device):
            _lib.check(lib.apx_ipattn_fwd(_lib.ptr(x), _lib.ptr(wq), _lib.ptr(cache), _lib.ptr(wo), _lib.ptr(wpack), _lib.ptr(bo),
                                          _lib.ptr(y), bc, s, c, heads, lt, la, float(proc.scale), _lib.dtype_code(dtype),
                                          _lib.ptr(work), wsb, _lib.stream_ptr()), "apx_ipattn_fwd")
        ctx_.save_for_backward(x, enc.detach().to(dtype).contiguous(), wq, wo, cache)
        ctx_.meta = (heads, lt, la, float(proc.scale), wk_ip.dtype, wv_ip.dtype)
        return y

    @staticmethod
    def backward(ctx_, dy):
        lib = _lib.load()
        x, enc, wq, wo, cache = ctx_.saved_tensors
        heads, lt, la, scale, kdt, vdt = ctx_.meta
        bc, s, c = x.shape
        dctx = enc.shape[-1]
        dy = dy.to(x.dtype).contiguous()
        dx = torch.empty_like(x)
        dwk = torch.empty((c, dctx), dtype=torch.float32, device=x.device)
        dwv = torch.empty((c, dctx), dtype=torch.float32, device=x.device)
        wsb = lib.apx_ipattn_bwd_workspace_bytes(bc, s, c, la)
        work = _lib.workspace(wsb, x.device, "ipattn_bwd")
        with torch.cuda.device(x.device):
            _lib.check(lib.apx_ipattn_bwd(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(wq), _lib.ptr(wo), _lib.ptr(cache), _lib.ptr(enc),
                                          _lib.ptr(dx), _lib.ptr(dwk), _lib.ptr(dwv), bc, s, c, heads, lt, la, dctx, scale,
                                          _lib.dtype_code(x.dtype), _lib.ptr(work), wsb, _lib.stream_ptr()), "apx_ipattn_bwd")
        return dx, None, dwk.to(kdt), dwv.to(vdt), None, None, None, None, None


class IPAttnProcessor2_0(nn.Module):
    r"""Decoupled (text + ``scale`` · audio) cross-attention processor on the B200 kernels.

    Reference: ``IPAttnProcessor2_0`` (attention_processor.py:297-470).  Same constructor:

    Args:
        hidden_size (`int`): hidden size of the attention layer (C).
        name (`str`): processor key, e.g. ``down_blocks.1.attentions.1.transformer_blocks.0.attn2.processor``.
        cross_attention_dim (`int`): channels of ``encoder_hidden_states`` (768).
        num_tokens (`int`, defaults to 4): number of leading TEXT tokens in the context; the rest is audio
            (attention_processor.py:400-403).
        scale (`float`, defaults to 1.0): ``ap_scale``, weight of the audio branch.
        do_copy (`bool`): initialise ``to_k_ip`` / ``to_v_ip`` from ``copied_cross_attention/<name>_{k,v}.bin``.
    """

    def __init__(self, hidden_size, name, cross_attention_dim=None, num_tokens=4, scale=1.0, do_copy=False):
        super().__init__()
        self.hidden_size = hidden_size
        self.cross_attention_dim = cross_attention_dim
        self.num_tokens = num_tokens
        self.scale = scale
        self.to_k_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)
        self.to_v_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)
        self.name = name
        self._ctx_cache = None  # (key, cache blob, keyed tensors)
        self._w_cache = None    # (key, (wq, wo, packed blob or None), keyed tensors)
        if do_copy:
            from .weights import load_copied_cross_attention
            k_weight, v_weight = load_copied_cross_attention(_copied_weight_dir(), name)
            self.to_k_ip.weight = nn.Parameter(k_weight.to(torch.float32))
            self.to_v_ip.weight = nn.Parameter(v_weight.to(torch.float32))
            self.to_k_ip.weight.requires_grad = True
            self.to_v_ip.weight.requires_grad = True

    # -- context K/V cache -------------------------------------------------------------------------------
    @staticmethod
    def _tkey(t: torch.Tensor):
        return (t.data_ptr(), t._version, tuple(t.shape), t.dtype)

    def invalidate_cache(self):
        """Drop every derived copy (context K/V blob, packed weights, folded-norm pack, fp32 bias): call after edits that do not bump
        ``_version`` (``param.data`` writes, raw kernels, graph replays into a keyed tensor)."""
        self._ctx_cache = None
        self._w_cache = None
        self._wln_cache = None
        self._bo_cache = None

    def _site_weights(self, attn, heads: int, dtype: torch.dtype):
        """(to_q.weight, to_out[0].weight, packed blob) in the kernel dtype; the blob is the one-time re-layout the
        fused kernel streams per head (apx_ipattn_pack), None when the fused kernel does not cover (C, heads).
        Cached per weight version: the UNet is frozen, so this runs once."""
        wq_p, wo_p = attn.to_q.weight, attn.to_out[0].weight
        key = (self._tkey(wq_p), self._tkey(wo_p), heads, dtype)
        if self._w_cache is not None and self._w_cache[0] == key:
            return self._w_cache[1]
        lib = _lib.load()
        c = self.hidden_size
        wq = wq_p.detach().to(dtype).contiguous()
        wo = wo_p.detach().to(dtype).contiguous()
        if tuple(wq.shape) != (c, c) or tuple(wo.shape) != (c, c):
            raise _lib.ApxError(f"{self.name}: to_q / to_out weights {tuple(wq.shape)} / {tuple(wo.shape)} != ({c}, {c})")
        pack = None
        nbytes = lib.apx_ipattn_pack_bytes(c, heads)
        if nbytes:
            pack = torch.empty(nbytes, dtype=torch.uint8, device=wq.device)
            with torch.cuda.device(wq.device):
                _lib.check(lib.apx_ipattn_pack(_lib.ptr(wq), _lib.ptr(wo), _lib.ptr(pack), c, heads, _lib.dtype_code(dtype),
                                               _lib.stream_ptr()), "apx_ipattn_pack")
        self._w_cache = (key, (wq, wo, pack), (wq_p, wo_p))
        return self._w_cache[1]

    def _context_cache(self, attn, ctx: torch.Tensor, heads: int, dtype: torch.dtype, force: bool = False) -> torch.Tensor:
        """Packed K / V of the context (apx_ctx_kv), cached per (context, weights) identity + version.  In-place edits that do
        not bump ``_version`` (``param.data`` writes, raw kernels, graph replays) need ``invalidate_cache()``; ``force`` rebuilds."""
        wk, wv = attn.to_k.weight, attn.to_v.weight
        wk_ip, wv_ip = self.to_k_ip.weight, self.to_v_ip.weight
        key = (self._tkey(ctx), self._tkey(wk), self._tkey(wv), self._tkey(wk_ip), self._tkey(wv_ip), heads, self.num_tokens)
        if not force and self._ctx_cache is not None and self._ctx_cache[0] == key:
            return self._ctx_cache[1]
        lib = _lib.load()
        bc, l, dctx = ctx.shape
        lt = min(self.num_tokens, l)
        la = l - lt
        c = self.hidden_size
        ws = [w.detach().to(dtype).contiguous() for w in (wk, wv, wk_ip, wv_ip)]
        for w in ws:
            if tuple(w.shape) != (c, dctx):
                raise _lib.ApxError(f"{self.name}: K/V projection weight {tuple(w.shape)} != ({c}, {dctx})")
        ctx_c = ctx.detach().to(dtype).contiguous()
        nbytes = lib.apx_ctx_kv_bytes(bc, lt, la, c, heads)
        cache = torch.empty(nbytes, dtype=torch.uint8, device=ctx.device)
        wsb = lib.apx_ctx_kv_workspace_bytes(bc, lt, la, c)
        work = _lib.workspace(wsb, ctx.device, "ctx_kv")
        with torch.cuda.device(ctx.device):
            _lib.check(lib.apx_ctx_kv(_lib.ptr(ctx_c), *[_lib.ptr(w) for w in ws], _lib.ptr(cache), bc, lt, la, dctx, c, heads,
                                      _lib.dtype_code(dtype), _lib.ptr(work), wsb, _lib.stream_ptr()), "apx_ctx_kv")
        # the keyed tensors are kept alive with the entry so that their addresses cannot be recycled under the key
        self._ctx_cache = (key, cache, (ctx, wk, wv, wk_ip, wv_ip))
        return cache

    # -- norm2 folded into the site kernel ----------------------------------------------------------------
    fold_norm_enabled = os.environ.get("APX_LN_FOLD_SITE", "1") != "0"

    def can_fold_norm(self, attn, hidden_states, norm, encoder_hidden_states=None) -> bool:
        """True when this call can take the UN-normalised stream and run `norm` (the block's norm2) inside the site kernel
        (apx_ipattn_fwd_ln): inference on CUDA in 16 bit at a shape the all-heads site kernel runs."""
        if not (self.fold_norm_enabled and hidden_states.is_cuda and not torch.is_grad_enabled() and hidden_states.dim() == 3
                and hidden_states.dtype in (torch.float16, torch.bfloat16) and norm.elementwise_affine
                and encoder_hidden_states is not None and encoder_hidden_states.dim() == 3
                and attn.spatial_norm is None and attn.group_norm is None and not attn.norm_cross
                and getattr(attn.to_q, "bias", None) is None and hidden_states.shape[-1] == self.hidden_size):
            return False
        bc, s, c = hidden_states.shape
        lt = min(self.num_tokens, encoder_hidden_states.shape[1])
        la = encoder_hidden_states.shape[1] - lt
        return _lib.load().apx_ipattn_site_kernel(bc, s, c, attn.heads, lt, la) == _lib.APX_SITE_ALLHEADS

    def _site_weights_ln(self, attn, heads: int, dtype: torch.dtype, norm):
        """(packed blob of (Wq * gamma, Wo), c1, c2) for norm folded into the q projection, cached per weight version."""
        wq_p, wo_p = attn.to_q.weight, attn.to_out[0].weight
        src = (wq_p, wo_p, norm.weight, norm.bias)
        key = tuple(self._tkey(t) for t in src) + (heads, dtype)
        c = getattr(self, "_wln_cache", None)
        if c is not None and c[0] == key and all(a is b for a, b in zip(c[2], src)):
            return c[1]
        lib = _lib.load()
        ch = self.hidden_size
        wqg, c1, c2 = _lib.fold_layernorm(wq_p.detach().to(dtype), None, norm.weight.detach(), norm.bias.detach())
        wo = wo_p.detach().to(dtype).contiguous()
        pack = torch.empty(lib.apx_ipattn_pack_bytes(ch, heads), dtype=torch.uint8, device=wqg.device)
        with torch.cuda.device(wqg.device):
            _lib.check(lib.apx_ipattn_pack(_lib.ptr(wqg), _lib.ptr(wo), _lib.ptr(pack), ch, heads, _lib.dtype_code(dtype), _lib.stream_ptr()),
                       "apx_ipattn_pack")
        self._wln_cache = (key, (pack, c1, c2), src)
        return self._wln_cache[1]

    # -- forward -----------------------------------------------------------------------------------------
    accepts_residual = True   # __call__ takes `residual_inout` (the restated transformer block passes its residual stream)
    folds_in_place = True     # ... and adds the site's output into it in place

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0,
                 residual_inout=None, fold_norm=None):
        """`fold_norm`: the block's norm2, when `hidden_states` is the un-normalised stream (only after can_fold_norm said yes)."""
        if scale != 1.0:
            warnings.warn("`scale` of IPAttnProcessor should be set by `set_ip_adapter_scale`.")
        if attention_mask is not None:
            raise NotImplementedError(
                "IPAttnProcessor2_0: attention_mask on the adapter site is not implemented (the AP-adapter pipeline never "
                "passes one for the 768-d context; reference attention_processor.py:378-382,424-428 is a dead path)")
        _lib.require_cuda(hidden_states, encoder_hidden_states)
        needs_grad = torch.is_grad_enabled() and (hidden_states.requires_grad or self.to_k_ip.weight.requires_grad
                                                  or self.to_v_ip.weight.requires_grad)
        residual = hidden_states
        if attn.spatial_norm is not None:
            hidden_states = attn.spatial_norm(hidden_states, temb)
        input_ndim = hidden_states.ndim
        if input_ndim == 4:
            batch_size, channel, height, width = hidden_states.shape
            hidden_states = hidden_states.view(batch_size, channel, height * width).transpose(1, 2)
        if encoder_hidden_states is None:
            # reference :389-390 would attend to the hidden states themselves; with to_k_ip built for the
            # cross-attention width this only type-checks when cross_attention_dim == hidden_size.
            encoder_hidden_states = hidden_states
        elif encoder_hidden_states.dim() < 3:
            encoder_hidden_states = encoder_hidden_states.unsqueeze(0)
        if attn.group_norm is not None:
            hidden_states = attn.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        if attn.norm_cross:
            encoder_hidden_states = attn.norm_encoder_hidden_states(encoder_hidden_states)

        if hidden_states.dtype == torch.float32 and torch.is_autocast_enabled():
            # training runs under bf16 / fp16 autocast with fp32 LayerNorm outputs (train_apadapter_v2.py:548-552): the linears
            # of the reference processor would cast their inputs exactly like this
            hidden_states = hidden_states.to(torch.get_autocast_dtype('cuda'))
        dtype = hidden_states.dtype
        if dtype not in (torch.float16, torch.bfloat16):
            raise _lib.ApxError(f"IPAttnProcessor2_0: hidden_states dtype {dtype}; the kernels take float16 / bfloat16 "
                                "(the reference runs inference in fp16, inference.py:13)")
        batch_size = encoder_hidden_states.shape[0]  # reference :374-376
        x = hidden_states.contiguous()
        if x.shape[0] != batch_size:
            raise _lib.ApxError(f"{self.name}: hidden_states batch {x.shape[0]} != context batch {batch_size}")
        bc, s, c = x.shape
        if c != self.hidden_size:
            raise _lib.ApxError(f"{self.name}: hidden size {c} != {self.hidden_size}")
        heads = attn.heads
        lt = min(self.num_tokens, encoder_hidden_states.shape[1])
        la = encoder_hidden_states.shape[1] - lt

        lib = _lib.load()
        if getattr(attn.to_q, "bias", None) is not None:
            raise _lib.ApxError(f"{self.name}: to_q with bias is not implemented (attention_bias=False in AudioLDM2)")
        if needs_grad:
            if la == 0:
                raise _lib.ApxError(f"{self.name}: training needs audio tokens behind the {self.num_tokens} text tokens")
            y = _IPAttnSiteFn.apply(x, encoder_hidden_states, self.to_k_ip.weight, self.to_v_ip.weight, self, attn, heads, lt, la)
            hidden_states = attn.to_out[1](y)
            if input_ndim == 4:
                hidden_states = hidden_states.transpose(-1, -2).reshape(batch_size, channel, height, width)
            if attn.residual_connection:
                hidden_states = hidden_states + residual
            if attn.rescale_output_factor != 1.0:
                hidden_states = hidden_states / attn.rescale_output_factor
            return hidden_states if residual_inout is None else hidden_states + residual_inout
        cache = self._context_cache(attn, encoder_hidden_states, heads, dtype)
        bo = attn.to_out[0].bias
        if bo is not None:
            bo = self._bias_f32(bo)
        if fold_norm is not None:
            # norm2 inside the site kernel: X is the raw stream, the pack holds Wq * gamma, the kernel takes the row statistics
            pack, c1, c2 = self._site_weights_ln(attn, heads, dtype, fold_norm)
            fold = (residual_inout is not None and input_ndim == 3 and not attn.residual_connection and attn.rescale_output_factor == 1.0
                    and residual_inout.shape == x.shape and residual_inout.dtype == dtype and residual_inout.is_contiguous()
                    and residual_inout.device == x.device and not (attn.to_out[1].training and attn.to_out[1].p > 0))
            y = residual_inout if fold else torch.empty_like(x)
            with torch.cuda.device(x.device):
                _lib.check(lib.apx_ipattn_fwd_ln(_lib.ptr(x), _lib.ptr(cache), _lib.ptr(pack), _lib.ptr(bo), _lib.ptr(c1), _lib.ptr(c2),
                                                 float(fold_norm.eps), _lib.ptr(y), bc, s, c, heads, lt, la, float(self.scale),
                                                 _lib.dtype_code(dtype), 1 if fold else 0, _lib.stream_ptr()), "apx_ipattn_fwd_ln")
            if fold:
                return y
            hidden_states = attn.to_out[1](y)
            if attn.rescale_output_factor != 1.0:
                hidden_states = hidden_states / attn.rescale_output_factor
            return hidden_states if residual_inout is None else hidden_states + residual_inout
        wq, wo, wpack = self._site_weights(attn, heads, dtype)
        wsb = lib.apx_ipattn_workspace_bytes(bc, s, c)
        work = _lib.workspace(wsb, x.device, "ipattn")
        # The block's residual add (h + attn2(norm2(h))) folded into the kernel: the site's output boxes are TMA-reduce-added
        # into the residual stream (three-kernel shapes: 16-bit residual operand of the to_out GEMM), in place.
        fold = (residual_inout is not None and input_ndim == 3 and not attn.residual_connection
                and attn.rescale_output_factor == 1.0 and residual_inout.shape == x.shape and residual_inout.dtype == dtype
                and residual_inout.is_contiguous() and residual_inout.device == x.device
                and not (attn.to_out[1].training and attn.to_out[1].p > 0))
        y = residual_inout if fold else torch.empty_like(x)
        fn, fname = (lib.apx_ipattn_fwd_add, "apx_ipattn_fwd_add") if fold else (lib.apx_ipattn_fwd, "apx_ipattn_fwd")
        with torch.cuda.device(x.device):
            _lib.check(fn(_lib.ptr(x), _lib.ptr(wq), _lib.ptr(cache), _lib.ptr(wo), _lib.ptr(wpack), _lib.ptr(bo),
                          _lib.ptr(y), bc, s, c, heads, lt, la, float(self.scale), _lib.dtype_code(dtype),
                          _lib.ptr(work), wsb, _lib.stream_ptr()), fname)
        if fold:
            return y
        hidden_states = attn.to_out[1](y)  # dropout (p = 0)
        if input_ndim == 4:
            hidden_states = hidden_states.transpose(-1, -2).reshape(batch_size, channel, height, width)
        if attn.residual_connection:
            hidden_states = hidden_states + residual
        if attn.rescale_output_factor != 1.0:
            hidden_states = hidden_states / attn.rescale_output_factor
        return hidden_states if residual_inout is None else hidden_states + residual_inout

    def _bias_f32(self, bo: torch.Tensor) -> torch.Tensor:
        """fp32 copy of the out-projection bias (the epilogue adds it to the fp32 accumulator), cached per version."""
        key = self._tkey(bo)
        cached = getattr(self, "_bo_cache", None)
        if cached is None or cached[0] != key:
            cached = (key, bo.detach().float().contiguous(), bo)
            self._bo_cache = cached
        return cached[1]
