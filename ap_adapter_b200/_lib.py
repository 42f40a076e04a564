"""ctypes binding of libapx.so (the C ABI declared in include/apx.h).

PyTorch is plumbing here: it owns device memory and streams; every arithmetic step of the hot path runs in
the hand-written sm_100a kernels behind this ABI.  There is no CPU or eager fallback — if the library is
missing, or a tensor is not on a CUDA device, the call fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libapx.so")

APX_F16, APX_BF16, APX_F32 = 0, 1, 2
APX_EPI_NONE, APX_EPI_GELU, APX_EPI_GEGLU, APX_EPI_RES16 = 0, 1, 2, 4
APX_IMPL_TC, APX_IMPL_SIMT, APX_IMPL_TC_UNFUSED = 0, 1, 2
APX_SITE_AUTO, APX_SITE_ALLHEADS, APX_SITE_ONEHEAD, APX_SITE_ALLHEADS4 = 0, 1, 2, 3

# every symbol include/apx.h declares: (restype, argtypes)
_P, _I, _L, _F, _Z = c_void_p, c_int, c_int64, c_float, c_size_t


class VitBlock(Structure):
    _fields_ = [(n, c_void_p) for n in (
        "ln1_w", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b",
        "qkv_wg", "qkv_c1", "qkv_c2", "fc1_wg", "fc1_c1", "fc1_c2")]   # ABI v3: LayerNorm-folded projections (None = separate LayerNorm)


class VitWeights(Structure):
    _fields_ = [("depth", c_int), ("embed", c_int), ("heads", c_int), ("mlp", c_int), ("wdtype", c_int),
                ("patch_wT", c_void_p), ("patch_b", c_void_p), ("pos_embed", c_void_p), ("cls_token", c_void_p),
                ("norm_w", c_void_p), ("norm_b", c_void_p), ("blocks", POINTER(VitBlock))]


SYMBOLS = {
    "apx_last_error": (c_char_p, []),
    "apx_version": (_I, []),
    "apx_device_ok": (_I, []),
    "apx_set_impl": (_I, [_I]),
    "apx_get_impl": (_I, []),
    "apx_launch_count": (c_uint64, []),
    "apx_linear": (_I, [_P, _L, _P, _L, _P, _P, _L, _P, _L, _I, _I, _I, _I, _I, _I, _P]),
    "apx_linear_lnfold": (_I, [_P, _L, _P, _L, _P, _P, _F, _P, _L, _P, _L, _I, _I, _I, _I, _I, _I, _P]),
    "apx_attention": (_I, [_P, _L, _L, _P, _L, _L, _P, _L, _L, _P, _L, _L, _I, _I, _I, _I, _I, _I, _F, _F, _I, _P]),
    "apx_flash_attention": (_I, [_P, _L, _L, _P, _L, _L, _P, _L, _L, _P, _L, _L, _I, _I, _I, _I, _I, _I, _P]),
    "apx_ctx_kv_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "apx_ctx_kv_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "apx_ctx_kv": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "apx_ipattn_workspace_bytes": (_Z, [_I, _I, _I]),
    "apx_ipattn_pack_bytes": (_Z, [_I, _I]),
    "apx_ipattn_pack": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "apx_ipattn_fused_ok": (_I, [_I, _I, _I, _I]),
    "apx_ipattn_site_kernel": (_I, [_I, _I, _I, _I, _I, _I]),
    "apx_set_site_variant": (_I, [_I]),
    "apx_ipattn_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _I, _P, _Z, _P]),
    "apx_ipattn_fwd_add": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _I, _P, _Z, _P]),
    "apx_ipattn_fwd_ln": (_I, [_P, _P, _P, _P, _P, _P, _F, _P, _I, _I, _I, _I, _I, _I, _F, _I, _I, _P]),
    "apx_ipattn_bwd_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "apx_ipattn_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P, _Z, _P]),
    "apx_fbank_batch": (_I, [_P, _I, _I, _P, _P, _I, _P, _P]),
    "apx_resample": (_I, [_P, _I, _I, _P, _I, _I, _I, _P, _P]),
    "apx_groupnorm_nhwc": (_I, [_P, _P, _P, _P, _P, _L, _P, _I, _I, _I, _I, _F, _I, _I, _P]),
    "apx_add_bias_nhwc": (_I, [_P, _P, _P, _P, _L, _I, _F, _I, _P]),
    "apx_add_layernorm": (_I, [_P, _P, _P, _P, _I, _P, _P, _L, _I, _F, _I, _P]),
    "apx_audiomae_workspace_bytes": (_Z, [_I]),
    "apx_audiomae_fwd": (_I, [_P, POINTER(VitWeights), _P, _I, _I, _I, _P, _Z, _P]),
    "apx_audiomae_encode": (_I, [_P, POINTER(VitWeights), _P, _I, _P, _Z, _P]),
    "apx_patch_embed": (_I, [_P, _P, _P, _P, _P, _P, _I, _P]),
    "apx_layernorm": (_I, [_P, _P, _P, _P, _I, _I, _F, _I, _P]),
    "apx_pool_tokens": (_I, [_P, _P, _I, _I, _I, _P]),
    "apx_fbank": (_I, [_P, _I, _P, _P, _I, _P, _P]),
}

_lib = None


class ApxError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libapx.so (built in-tree by `__graft_entry__.build()` / `make -C ap_adapter_b200/csrc`)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ApxError(
                f"{LIB_PATH} not found: the sm_100a extension is not built (run `python -c 'import __graft_entry__ as g; "
                "g.build()'`). ap_adapter_b200 has no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError here == header / library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise ApxError(f"{what} failed (code {rc}): {load().apx_last_error().decode(errors='replace')}")


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float16:
        return APX_F16
    if dt == torch.bfloat16:
        return APX_BF16
    if dt == torch.float32:
        return APX_F32
    raise ApxError(f"unsupported dtype {dt}: the sm_100a kernels take float16 / bfloat16 operands")


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise ApxError("ap_adapter_b200 runs on CUDA tensors only (no CPU fallback); got a tensor on " + str(t.device))
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ApxError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


def ptr(t: torch.Tensor | None):
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


_workspaces: dict[tuple, torch.Tensor] = {}
_ws_scope: dict | None = None   # set by workspace_scope(): scratch buffers owned by one CUDA graph


class workspace_scope:
    """While active, ``workspace()`` hands out buffers from ``owner`` (a dict kept by the caller) instead of the process-wide
    per-stream table.  A CUDA-graph capture runs inside one: the scratch buffers are then allocated in THAT graph's pool, live
    exactly as long as the graph state that owns the dict, and two graphs never bake in the same scratch pointer."""

    def __init__(self, owner: dict):
        self.owner = owner

    def __enter__(self):
        global _ws_scope
        self.prev, _ws_scope = _ws_scope, self.owner
        return self.owner

    def __exit__(self, *exc):
        global _ws_scope
        _ws_scope = self.prev


def workspace(nbytes: int, device: torch.device, tag: str = "") -> torch.Tensor:
    """A grow-only scratch buffer per (device, stream, tag) — or per (device, tag) of the active ``workspace_scope``.  Kernels
    on one stream are serialised, so one buffer per stream is race-free (one host thread per stream is assumed); nothing is
    allocated on the steady-state path (CUDA-graph friendly)."""
    if _ws_scope is not None:
        table, key = _ws_scope, (device.index, tag)
    else:
        table, key = _workspaces, (device.index, torch.cuda.current_stream(device).cuda_stream, tag)
    buf = table.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        table[key] = buf
    return buf


def launch_count() -> int:
    return int(load().apx_launch_count())


# ---- thin functional wrappers (used by the processors, the encoder and the tests) ----------------------
def linear(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, residual: torch.Tensor | None = None,
           out_dtype: torch.dtype | None = None, gelu: bool = False, out: torch.Tensor | None = None,
           geglu: bool = False) -> torch.Tensor:
    """out[M,N] = act(a[M,K] @ w[N,K]^T + bias) + residual   (apx_linear).

    geglu=True: rows of `w` are interleaved (hidden_0, gate_0, hidden_1, gate_1, ...) and the result has N/2 columns,
    hidden * gelu(gate).  `residual` is fp32, or the output dtype (16-bit) — the flag APX_EPI_RES16 is derived from it."""
    lib = load()
    dev = require_cuda(a, w, bias, residual, out)
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1] and a.dtype == w.dtype
    assert not (gelu and geglu)
    a = a if a.stride(1) == 1 else a.contiguous()
    w = w if w.stride(1) == 1 else w.contiguous()
    m, k = a.shape
    n = w.shape[0]
    out_dtype = out_dtype or a.dtype
    if out is None:
        out = torch.empty((m, n // 2 if geglu else n), dtype=out_dtype, device=dev)
    epi = APX_EPI_GEGLU if geglu else APX_EPI_GELU if gelu else APX_EPI_NONE
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
    if residual is not None:
        assert residual.stride(1) == 1 and residual.shape == out.shape
        if residual.dtype != torch.float32:
            assert residual.dtype == out.dtype and out.dtype != torch.float32
            epi |= APX_EPI_RES16
    with torch.cuda.device(dev):
        check(lib.apx_linear(ptr(a), a.stride(0), ptr(w), w.stride(0), ptr(bias), ptr(residual),
                             residual.stride(0) if residual is not None else 0, ptr(out), out.stride(0), m, n, k,
                             dtype_code(a.dtype), dtype_code(out.dtype), epi, stream_ptr()), "apx_linear")
    return out


def fold_layernorm(w: torch.Tensor, bias: torch.Tensor | None, gamma: torch.Tensor, beta: torch.Tensor):
    """(Wg, c1, c2) of apx_linear_lnfold for a projection ``w`` [N, K] (+ ``bias``) applied to LayerNorm(x; gamma, beta):
    Wg = w * gamma in w's dtype, c1 = row sums of the ROUNDED Wg, c2 = w @ beta + bias (fp32)."""
    with torch.no_grad():
        wg = (w.float() * gamma.float()[None, :]).to(w.dtype).contiguous()
        c1 = wg.float().sum(dim=1).contiguous()
        c2 = w.float() @ beta.float()
        if bias is not None:
            c2 = c2 + bias.float()
        return wg, c1, c2.contiguous()


def linear_lnfold(a: torch.Tensor, wg: torch.Tensor, c1: torch.Tensor, c2: torch.Tensor, eps: float, residual: torch.Tensor | None = None,
                  out_dtype: torch.dtype | None = None, gelu: bool = False, geglu: bool = False) -> torch.Tensor:
    """act(LayerNorm(a) @ w^T + b) + residual with the LayerNorm folded into the GEMM (apx_linear_lnfold; see fold_layernorm)."""
    lib = load()
    dev = require_cuda(a, wg, c1, c2, residual)
    assert a.dim() == 2 and wg.dim() == 2 and a.shape[1] == wg.shape[1] and a.dtype == wg.dtype and not (gelu and geglu)
    a = a if a.stride(1) == 1 else a.contiguous()
    m, k = a.shape
    n = wg.shape[0]
    out_dtype = out_dtype or a.dtype
    out = torch.empty((m, n // 2 if geglu else n), dtype=out_dtype, device=dev)
    epi = APX_EPI_GEGLU if geglu else APX_EPI_GELU if gelu else APX_EPI_NONE
    assert c1.dtype == torch.float32 and c2.dtype == torch.float32 and c1.numel() == n and c2.numel() == n
    if residual is not None:
        assert residual.stride(1) == 1 and residual.shape == out.shape
        if residual.dtype != torch.float32:
            assert residual.dtype == out.dtype and out.dtype != torch.float32
            epi |= APX_EPI_RES16
    with torch.cuda.device(dev):
        check(lib.apx_linear_lnfold(ptr(a), a.stride(0), ptr(wg), wg.stride(0), ptr(c1), ptr(c2), float(eps), ptr(residual),
                                    residual.stride(0) if residual is not None else 0, ptr(out), out.stride(0), m, n, k,
                                    dtype_code(a.dtype), dtype_code(out.dtype), epi, stream_ptr()), "apx_linear_lnfold")
    return out


def flash_attention_ok(channels: int, heads: int) -> bool:
    """Head sizes apx_flash_attention covers (d % 16 == 0, 16 <= d <= 80)."""
    d = channels // heads if heads > 0 and channels % heads == 0 else 0
    return d % 16 == 0 and 16 <= d <= 80


def flash_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """softmax(q k^T / sqrt(d)) v on token-major (B, S, H*d) / (B, L, H*d) views, any number of keys (apx_flash_attention)."""
    lib = load()
    dev = require_cuda(q, k, v)
    assert q.dim() == 3 and k.dim() == 3 and v.dim() == 3 and k.shape == v.shape and q.shape[0] == k.shape[0] and q.shape[2] == k.shape[2]
    assert q.dtype == k.dtype == v.dtype and q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1
    b, s, c = q.shape
    out = torch.empty((b, s, c), dtype=q.dtype, device=dev)
    with torch.cuda.device(dev):
        check(lib.apx_flash_attention(ptr(q), q.stride(0), q.stride(1), ptr(k), k.stride(0), k.stride(1), ptr(v), v.stride(0),
                                      v.stride(1), ptr(out), out.stride(0), out.stride(1), b, s, k.shape[1], c, heads,
                                      dtype_code(q.dtype), stream_ptr()), "apx_flash_attention")
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, l0: int | None = None,
              w0: float = 1.0, w1: float = 0.0) -> torch.Tensor:
    """Segmented SDPA on token-major (B, S, H*d) / (B, L, H*d) views (apx_attention)."""
    lib = load()
    dev = require_cuda(q, k, v)
    b, s, c = q.shape
    l = k.shape[1]
    l0 = l if l0 is None else l0
    for t in (q, k, v):
        assert t.stride(2) == 1 and t.dtype == q.dtype
    o = torch.empty((b, s, c), dtype=q.dtype, device=dev)
    with torch.cuda.device(dev):
        check(lib.apx_attention(ptr(q), q.stride(0), q.stride(1), ptr(k), k.stride(0), k.stride(1), ptr(v), v.stride(0),
                                v.stride(1), ptr(o), o.stride(0), o.stride(1), b, s, heads, c // heads, l, l0, w0, w1,
                                dtype_code(q.dtype), stream_ptr()), "apx_attention")
    return o


def groupnorm_nhwc_ok(hw: int, c: int, groups: int) -> bool:
    """Shapes apx_groupnorm_nhwc takes (one group's HW x C/groups values are staged as fp32 in the shared memory of a cluster of up to
    8 CTAs)."""
    return c % groups == 0 and (c // groups) % 4 == 0 and hw * (c // groups) * 4 <= 8 * 190 * 1024


def groupnorm_nhwc(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float, silu: bool,
                   chan_bias: torch.Tensor | None = None, bc_add: torch.Tensor | None = None) -> torch.Tensor:
    """[SiLU] GroupNorm(x + chan_bias[c] + bc_add[b, c]) on a channels-last (B, C, H, W) 16-bit tensor (apx_groupnorm_nhwc).
    gamma / beta / chan_bias fp32 (C,), bc_add (B, C) in x's dtype.  Returns a channels-last tensor of x's shape."""
    lib = load()
    dev = require_cuda(x, gamma, beta, chan_bias, bc_add)
    b, c, h, w = x.shape
    assert x.is_contiguous(memory_format=torch.channels_last) and gamma.dtype == torch.float32 and beta.dtype == torch.float32
    y = torch.empty_like(x, memory_format=torch.channels_last)
    if bc_add is not None:
        assert bc_add.shape == (b, c) and bc_add.dtype == x.dtype and bc_add.stride(1) == 1
    if chan_bias is not None:
        assert chan_bias.dtype == torch.float32 and chan_bias.numel() == c
    with torch.cuda.device(dev):
        check(lib.apx_groupnorm_nhwc(ptr(x), ptr(gamma), ptr(beta), ptr(chan_bias), ptr(bc_add), bc_add.stride(0) if bc_add is not None else 0,
                                     ptr(y), b, h * w, c, groups, float(eps),
                                     1 if silu else 0, dtype_code(x.dtype), stream_ptr()), "apx_groupnorm_nhwc")
    return y


def add_bias_nhwc(a: torch.Tensor, b: torch.Tensor, bias: torch.Tensor | None, scale: float = 1.0) -> torch.Tensor:
    """(a + b + bias[c]) * scale on channels-last (B, C, H, W) 16-bit tensors (apx_add_bias_nhwc)."""
    lib = load()
    dev = require_cuda(a, b, bias)
    assert a.shape == b.shape and a.dtype == b.dtype
    assert a.is_contiguous(memory_format=torch.channels_last) and b.is_contiguous(memory_format=torch.channels_last)
    n, c, h, w = a.shape
    y = torch.empty_like(a, memory_format=torch.channels_last)
    with torch.cuda.device(dev):
        check(lib.apx_add_bias_nhwc(ptr(a), ptr(b), ptr(bias), ptr(y), n * h * w, c, float(scale), dtype_code(a.dtype), stream_ptr()),
              "apx_add_bias_nhwc")
    return y


def add_layernorm_ok(x: torch.Tensor, norm) -> bool:
    """Shapes / modules apx_add_layernorm takes: a contiguous 16-bit (..., C) tensor, C % 8 == 0, C <= 2048, an affine LayerNorm over C."""
    w = getattr(norm, "weight", None)
    return (x.is_cuda and x.dtype in (torch.float16, torch.bfloat16) and x.is_contiguous() and x.shape[-1] % 8 == 0 and x.shape[-1] <= 2048
            and w is not None and norm.bias is not None and tuple(norm.normalized_shape) == (x.shape[-1],)
            and w.dtype in (x.dtype, torch.float32) and norm.bias.dtype == w.dtype and w.is_contiguous() and norm.bias.is_contiguous())


def add_layernorm(a: torch.Tensor, b: torch.Tensor | None, norm) -> tuple[torch.Tensor | None, torch.Tensor]:
    """(a + b, LayerNorm(a + b)) in one kernel (apx_add_layernorm); b None: (None, LayerNorm(a)).  `norm` is the nn.LayerNorm."""
    lib = load()
    w, bias = norm.weight.detach(), norm.bias.detach()
    dev = require_cuda(a, b, w, bias)
    assert b is None or (b.shape == a.shape and b.dtype == a.dtype and b.is_contiguous())
    c = a.shape[-1]
    y_ln = torch.empty_like(a)
    y_sum = torch.empty_like(a) if b is not None else None
    with torch.cuda.device(dev):
        check(lib.apx_add_layernorm(ptr(a), ptr(b), ptr(w), ptr(bias), dtype_code(w.dtype), ptr(y_sum), ptr(y_ln), a.numel() // c, c,
                                    float(norm.eps), dtype_code(a.dtype), stream_ptr()), "apx_add_layernorm")
    return y_sum, y_ln


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float, out_dtype: torch.dtype) -> torch.Tensor:
    lib = load()
    dev = require_cuda(x, w, b)
    assert x.dtype == torch.float32 and x.is_contiguous() and w.dtype == torch.float32 and b.dtype == torch.float32
    rows, cols = x.numel() // x.shape[-1], x.shape[-1]
    y = torch.empty(x.shape, dtype=out_dtype, device=dev)
    with torch.cuda.device(dev):
        check(lib.apx_layernorm(ptr(x), ptr(w), ptr(b), ptr(y), rows, cols, eps, dtype_code(out_dtype), stream_ptr()),
              "apx_layernorm")
    return y
