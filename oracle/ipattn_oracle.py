"""fp32 CPU restatement of the reference attention processors (TEST INFRASTRUCTURE — see oracle/__init__.py).

Functional forms: weights are passed as tensors so the same function serves the golden-vector check, the GPU
parity tests and the CPU baseline.  Everything is computed in the dtype of the inputs (fp32 on the CPU).
"""
from __future__ import annotations

import math

import torch


def _split_heads(t: torch.Tensor, heads: int) -> torch.Tensor:
    b, n, c = t.shape
    return t.view(b, n, heads, c // heads).transpose(1, 2)  # (B, H, N, d)


def _sdpa(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
    """softmax(q k^T / sqrt(d) + mask) v, restating F.scaled_dot_product_attention's default scale.

    An empty key set yields zeros, as torch's SDPA does (attention_processor.py:443 with no audio tokens).
    """
    if k.shape[-2] == 0:
        return torch.zeros_like(q)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    if mask is not None:
        s = s + mask
    return torch.matmul(torch.softmax(s, dim=-1), v)


def ipattn_forward(
    hidden_states: torch.Tensor,          # (Bc, S, C)
    encoder_hidden_states: torch.Tensor,  # (Bc, Lt+La, Dctx) or (Lt+La, Dctx)
    wq: torch.Tensor,                     # (C, C)        attn.to_q.weight
    wk: torch.Tensor,                     # (C, Dctx)     attn.to_k.weight
    wv: torch.Tensor,                     # (C, Dctx)     attn.to_v.weight
    wk_ip: torch.Tensor,                  # (C, Dctx)     processor.to_k_ip.weight
    wv_ip: torch.Tensor,                  # (C, Dctx)     processor.to_v_ip.weight
    wo: torch.Tensor,                     # (C, C)        attn.to_out[0].weight
    bo: torch.Tensor | None,              # (C,)          attn.to_out[0].bias
    heads: int,
    num_tokens: int,
    scale: float,
    rescale_output_factor: float = 1.0,
) -> torch.Tensor:
    """IPAttnProcessor2_0.__call__ on the path the pipeline takes (attention_processor.py:347-470):
    3-D hidden states, no mask, no spatial/group/cross norm, no residual connection.

    :371-372 2-D context is unsqueezed; :387 to_q; :400-403 context split (text FIRST, audio after);
    :406-415 text K/V; :429-434 SDPA + head merge; :435-451 audio K/V from to_k_ip/to_v_ip, unmasked SDPA;
    :454 h + scale * h_ip; :457 to_out[0]; :468 / rescale_output_factor.
    """
    if encoder_hidden_states.dim() < 3:
        encoder_hidden_states = encoder_hidden_states.unsqueeze(0)
    batch = encoder_hidden_states.shape[0]  # :374-376 batch is taken from the context
    q = hidden_states @ wq.t()
    text = encoder_hidden_states[:, :num_tokens, :]
    audio = encoder_hidden_states[:, num_tokens:, :]
    c = q.shape[-1]

    qh = q.view(batch, -1, heads, c // heads).transpose(1, 2)
    h_text = _sdpa(qh, _split_heads(text @ wk.t(), heads), _split_heads(text @ wv.t(), heads))
    h_text = h_text.transpose(1, 2).reshape(batch, -1, c)
    h_ip = _sdpa(qh, _split_heads(audio @ wk_ip.t(), heads), _split_heads(audio @ wv_ip.t(), heads))
    h_ip = h_ip.transpose(1, 2).reshape(batch, -1, c)

    out = h_text + scale * h_ip
    out = out @ wo.t()
    if bo is not None:
        out = out + bo
    return out / rescale_output_factor


def attn_forward(
    hidden_states: torch.Tensor,                 # (B, S, C)
    encoder_hidden_states: torch.Tensor | None,  # (B, L, Dctx) or None (self-attention)
    wq: torch.Tensor, wk: torch.Tensor, wv: torch.Tensor, wo: torch.Tensor, bo: torch.Tensor | None,
    heads: int,
    attention_mask: torch.Tensor | None = None,  # additive, broadcastable to (B, H, S, L)
) -> torch.Tensor:
    """AttnProcessor2_0.__call__ (attention_processor.py:214-294) on the 3-D / no-norm path."""
    ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
    if ctx.dim() < 3:
        ctx = ctx.unsqueeze(0)
    b = ctx.shape[0]
    c = wq.shape[0]
    q = (hidden_states @ wq.t()).view(b, -1, heads, c // heads).transpose(1, 2)
    k = _split_heads(ctx @ wk.t(), heads)
    v = _split_heads(ctx @ wv.t(), heads)
    h = _sdpa(q, k, v, attention_mask).transpose(1, 2).reshape(b, -1, c)
    out = h @ wo.t()
    if bo is not None:
        out = out + bo
    return out
