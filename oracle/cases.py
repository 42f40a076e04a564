"""Seeded parity cases shared by oracle/gen_golden.py and tests/ (TEST INFRASTRUCTURE).

Inputs are drawn from numpy's frozen legacy ``RandomState`` so they are bit-identical on any machine and only
the reference's OUTPUTS have to be stored under tests/golden/.
"""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
DCTX = 768

# name -> (Bc, S, C, H, Lt(num_tokens), La, ap_scale, ctx_2d, real_ip_weights: False or the scale applied to them)
IPATTN_CASES = {
    # BASELINE.json config 1: Q(2,1024,320), text-KV(2,77,768), audio-KV(2,8,768), ap_scale 0.5
    "cfg1":        (2, 1024, 320, 8, 77, 8, 0.5, False, False),
    # real site geometries, 10 s latent => S = 1000 / 252 / 64
    "site256":     (2, 1000, 256, 8, 8, 128, 0.5, False, 8.0),    # shipped adapter-init weights x8 (sharp audio softmax)
    "site256_real": (2, 1000, 256, 8, 8, 128, 0.5, False, 1.0),   # the same weights exactly as shipped
    "site384":     (2, 252, 384, 8, 8, 32, 0.55, False, False),
    "site640":     (2, 64, 640, 8, 8, 8, 0.5, False, False),
    "site256_la512": (1, 96, 256, 8, 8, 512, 1.0, False, False),
    "ctx2d":       (1, 40, 256, 8, 8, 32, 0.5, True, False),
    "text_only":   (2, 33, 256, 8, 8, 0, 0.5, False, False),
}
# rows of the (Bc*S, C) output kept in the fixture (every k-th) to keep the files small
ROW_STRIDE = {"cfg1": 8, "site256": 8, "site256_real": 8}

REAL_SITE = "down_blocks.1.attentions.1.transformer_blocks.0.attn2.processor"


def _randn(rs: np.random.RandomState, *shape, std=1.0) -> torch.Tensor:
    return torch.from_numpy((rs.standard_normal(shape) * std).astype(np.float32))


def make_ipattn_case(name: str) -> dict:
    """Returns dict(x, ctx, wq, wk, wv, wk_ip, wv_ip, wo, bo, heads, num_tokens, scale) as fp32 CPU tensors."""
    bc, s, c, h, lt, la, scale, ctx2d, real = IPATTN_CASES[name]
    rs = np.random.RandomState(abs(hash_name(name)) % (2 ** 31))
    x = _randn(rs, bc, s, c)
    ctx = _randn(rs, bc, lt + la, DCTX)
    if ctx2d:
        ctx = ctx[0]
    case = dict(
        x=x, ctx=ctx,
        wq=_randn(rs, c, c, std=c ** -0.5), wk=_randn(rs, c, DCTX, std=1.5 * DCTX ** -0.5),
        wv=_randn(rs, c, DCTX, std=DCTX ** -0.5), wk_ip=_randn(rs, c, DCTX, std=1.5 * DCTX ** -0.5),
        wv_ip=_randn(rs, c, DCTX, std=DCTX ** -0.5), wo=_randn(rs, c, c, std=c ** -0.5), bo=_randn(rs, c, std=0.1),
        heads=h, num_tokens=lt, scale=scale,
    )
    if real:
        # the shipped adapter-init weights (reference copied_cross_attention/*.bin), stored fp16 in the fixture
        f = np.load(os.path.join(GOLDEN_DIR, "real_ip_weights.npz"))
        case["wk_ip"] = torch.from_numpy(f["k"].astype(np.float32)) * float(real)   # x8: a sharp softmax; x1: as shipped
        case["wv_ip"] = torch.from_numpy(f["v"].astype(np.float32)) * float(real)
    return case


def hash_name(name: str) -> int:
    """Stable (non-salted) string hash for seeding."""
    hv = 2166136261
    for ch in name.encode():
        hv = ((hv ^ ch) * 16777619) & 0xFFFFFFFF
    return hv


# name -> (B, L, heads, C, Dctx or None for self-attention, masked)
ATTN_CASES = {
    "self_256": (2, 60, 8, 256, None, False),
    "t5_cross_384": (2, 48, 8, 384, 1024, True),
}


def make_attn_case(name: str) -> dict:
    b, s, h, c, dctx, masked = ATTN_CASES[name]
    rs = np.random.RandomState(hash_name(name) % (2 ** 31))
    x = _randn(rs, b, s, c)
    din = dctx or c
    case = dict(x=x, ctx=None, mask=None, heads=h,
                wq=_randn(rs, c, c, std=c ** -0.5), wk=_randn(rs, c, din, std=1.5 * din ** -0.5),
                wv=_randn(rs, c, din, std=din ** -0.5), wo=_randn(rs, c, c, std=c ** -0.5), bo=_randn(rs, c, std=0.1))
    if dctx:
        lctx = 16
        case["ctx"] = _randn(rs, b, lctx, dctx)
        if masked:
            # the UNet turns the (B, L) keep-mask into an additive (B, 1, L) bias: (1 - m) * -10000
            # (reference pipeline/modeling_audioldm2.py:741-747)
            keep = torch.ones(b, lctx)
            keep[0, 11:] = 0
            keep[1, 5:] = 0
            case["mask"] = ((1 - keep) * -10000.0).unsqueeze(1)
    return case


# ---- AudioMAE ----------------------------------------------------------------------------------
MAE_SEED = 7
POOLS = [(1, 1), (2, 2), (4, 4), (8, 8), (2, 1), (8, 2)]


def make_mel(seed: int = 3, batch: int = 1) -> torch.Tensor:
    """Synthetic log-mel: N(-0.45, 0.45^2), the statistics of a normalised Kaldi fbank."""
    rs = np.random.RandomState(seed)
    return torch.from_numpy((rs.standard_normal((batch, 1024, 128)) * 0.45 - 0.45).astype(np.float32))


def mae_state_dict(seed: int = MAE_SEED, depth: int = 12) -> dict:
    """Random AudioMAE encoder weights under the reference's key names, numpy-seeded (no checkpoint offline)."""
    from .audiomae_oracle import EMBED, MLP, sincos_pos_embed
    rs = np.random.RandomState(seed)
    n = lambda *s, std=0.02: _randn(rs, *s, std=std)  # noqa: E731
    sd = {
        "patch_embed.proj.weight": n(EMBED, 1, 16, 16, std=0.05),
        "patch_embed.proj.bias": n(EMBED),
        "cls_token": n(1, 1, EMBED),
        "pos_embed": sincos_pos_embed(),
        "norm.weight": 1.0 + n(EMBED, std=0.1),
        "norm.bias": n(EMBED, std=0.1),
    }
    for i in range(depth):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = 1.0 + n(EMBED, std=0.1)
        sd[p + "norm1.bias"] = n(EMBED, std=0.1)
        sd[p + "attn.qkv.weight"] = n(3 * EMBED, EMBED, std=0.04)
        sd[p + "attn.qkv.bias"] = n(3 * EMBED)
        sd[p + "attn.proj.weight"] = n(EMBED, EMBED)
        sd[p + "attn.proj.bias"] = n(EMBED)
        sd[p + "norm2.weight"] = 1.0 + n(EMBED, std=0.1)
        sd[p + "norm2.bias"] = n(EMBED, std=0.1)
        sd[p + "mlp.fc1.weight"] = n(MLP, EMBED)
        sd[p + "mlp.fc1.bias"] = n(MLP)
        sd[p + "mlp.fc2.weight"] = n(EMBED, MLP)
        sd[p + "mlp.fc2.bias"] = n(EMBED)
    return sd
