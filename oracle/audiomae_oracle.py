"""fp32 CPU restatement of the AudioMAE encoder path (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows reference ``audio_encoder/models_mae.py`` (PatchEmbed_org :19-44, sin-cos table :87-151,
``forward_encoder_no_random_mask_no_average`` :548-570, ``mae_vit_base_patch16`` :689-701) and
``audio_encoder/AudioMAE.py`` (``pool`` :148-182, ``forward`` :190-212).  The timm ``Block`` the reference
imports (``models_mae.py:18``) is NOT vendored there (timm is an unpinned requirement, ``requirements.txt:13``);
its published forward is restated here: pre-LN, fused qkv Linear, softmax((q * d^-0.5) k^T) v, proj, and an
erf-GELU MLP, no LayerScale / DropPath at the defaults the reference uses.

Weights are a ``state_dict`` with the reference's own key names (``patch_embed.proj.weight``, ``cls_token``,
``pos_embed``, ``blocks.N.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}.{weight,bias}``, ``norm.*``).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

EMBED, DEPTH, HEADS, MLP = 768, 12, 12, 3072
LN_EPS = 1e-6  # partial(nn.LayerNorm, eps=1e-6), models_mae.py:698


def _sincos_1d(embed_dim: int, pos: np.ndarray) -> np.ndarray:
    """models_mae.py:108-127."""
    omega = np.arange(embed_dim // 2, dtype=float)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_pos_embed(embed_dim: int = EMBED, grid_size=(8, 64)) -> torch.Tensor:
    """get_2d_sincos_pos_embed_flexible(embed_dim, patch_hw, cls_token=True)  (models_mae.py:136-151,330-340).

    ``patch_hw`` is (freq patches, time patches) = (8, 64) (models_mae.py:27) although tokens are ordered
    time-major — the table is reproduced as the reference builds it, not as it "should" be.
    """
    grid_h = np.arange(grid_size[0], dtype=np.float32)
    grid_w = np.arange(grid_size[1], dtype=np.float32)
    grid = np.stack(np.meshgrid(grid_w, grid_h), axis=0).reshape([2, 1, grid_size[0], grid_size[1]])
    emb = np.concatenate([_sincos_1d(embed_dim // 2, grid[0]), _sincos_1d(embed_dim // 2, grid[1])], axis=1)
    emb = np.concatenate([np.zeros([1, embed_dim]), emb], axis=0)
    return torch.from_numpy(emb).float().unsqueeze(0)  # (1, 513, 768)


def patch_embed(mel: torch.Tensor, sd: dict) -> torch.Tensor:
    """(B,1,1024,128) -> tokens (B,513,768): conv16x16/s16, +pos[1:], prepend cls+pos[0] (models_mae.py:42-43,550-563)."""
    x = F.conv2d(mel, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=16)
    x = x.flatten(2).transpose(1, 2)
    x = x + sd["pos_embed"][:, 1:, :]
    cls = (sd["cls_token"] + sd["pos_embed"][:, :1, :]).expand(x.shape[0], -1, -1)
    return torch.cat((cls, x), dim=1)


def vit_block(x: torch.Tensor, sd: dict, i: int, heads: int = HEADS) -> torch.Tensor:
    p = f"blocks.{i}."
    b, n, c = x.shape
    d = c // heads
    h = F.layer_norm(x, (c,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LN_EPS)
    qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(b, n, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = torch.softmax((q * d ** -0.5) @ k.transpose(-2, -1), dim=-1)
    h = (attn @ v).transpose(1, 2).reshape(b, n, c)
    x = x + F.linear(h, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    h = F.layer_norm(x, (c,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], LN_EPS)
    h = F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))  # exact erf GELU
    return x + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])


def encoder(mel: torch.Tensor, sd: dict, depth: int | None = None) -> torch.Tensor:
    """forward_encoder_no_random_mask_no_average (models_mae.py:548-570): (B,1,1024,128) -> (B,513,768)."""
    if depth is None:
        depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    x = patch_embed(mel, sd)
    for i in range(depth):
        x = vit_block(x, sd, i)
    return F.layer_norm(x, (x.shape[-1],), sd["norm.weight"], sd["norm.bias"], LN_EPS)


def pool(representation: torch.Tensor, time_pool: int, freq_pool: int) -> torch.Tensor:
    """AudioMAEConditionCTPoolRand.pool (AudioMAE.py:148-182): drop CLS, (B,768,64,8), (avg+max)/2."""
    assert representation.size(-1) == 768
    r = representation[:, 1:, :].transpose(1, 2)
    bs, dim, _ = r.shape
    r = r.reshape(bs, dim, 64, 8)
    k = (time_pool, freq_pool)
    pooled = (F.avg_pool2d(r, k, k) + F.max_pool2d(r, k, k)) / 2
    return pooled.flatten(2).transpose(1, 2)


def audiomae_forward(mel: torch.Tensor, sd: dict, time_pool: int, freq_pool: int, depth: int | None = None):
    """AudioMAEConditionCTPoolRand.forward (AudioMAE.py:190-212): mel (B,1024,128) -> [emb (B,L,768), ones (B,L)]."""
    assert mel.size(-2) == 1024 and mel.size(-1) == 128
    with torch.no_grad():
        rep = pool(encoder(mel.unsqueeze(1), sd, depth), time_pool, freq_pool)
    return [rep, torch.ones((rep.size(0), rep.size(1))).float()]


__all__ = ["sincos_pos_embed", "patch_embed", "vit_block", "encoder", "pool", "audiomae_forward",
           "EMBED", "DEPTH", "HEADS", "MLP", "LN_EPS"]
