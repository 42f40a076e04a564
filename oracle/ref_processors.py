"""nn.Module wrappers around the oracle functions with the reference's processor interface (TEST INFRASTRUCTURE).

``OracleIPAttnProcessor2_0`` has the constructor / ``__call__`` signature of the reference's ``IPAttnProcessor2_0``
(``APadapter/ap_adapter/attention_processor.py:297-470``) and computes with ``ipattn_oracle.ipattn_forward`` (plain
torch ops, fp32 on the CPU).  It is what the CPU legs of ``bench.py`` (``cpu_baseline`` / ``--impl reference``) and the
UNet-level parity tests install at the 32 adapter sites.  Never imported by ``ap_adapter_b200``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ipattn_oracle


class OracleIPAttnProcessor2_0(nn.Module):
    def __init__(self, hidden_size, name, cross_attention_dim=None, num_tokens=4, scale=1.0, do_copy=False):
        super().__init__()
        self.hidden_size, self.cross_attention_dim, self.num_tokens, self.scale, self.name = (
            hidden_size, cross_attention_dim, num_tokens, scale, name)
        self.to_k_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)
        self.to_v_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0):
        assert attention_mask is None, "the 768-d context is never masked on this path (pipeline_audioldm2.py:1010-1017)"
        return ipattn_oracle.ipattn_forward(
            hidden_states, encoder_hidden_states, attn.to_q.weight, attn.to_k.weight, attn.to_v.weight,
            self.to_k_ip.weight, self.to_v_ip.weight, attn.to_out[0].weight, attn.to_out[0].bias, attn.heads,
            self.num_tokens, self.scale, attn.rescale_output_factor)


class OracleAttnProcessor2_0(nn.Module):
    """Reference ``AttnProcessor2_0`` (attention_processor.py:199-294) on the 3-D / no-norm path."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        if attention_mask is not None:  # (B, 1, L) additive bias -> broadcast over heads and queries
            attention_mask = attention_mask.unsqueeze(1)
        return ipattn_oracle.attn_forward(hidden_states, encoder_hidden_states, attn.to_q.weight, attn.to_k.weight,
                                          attn.to_v.weight, attn.to_out[0].weight, attn.to_out[0].bias, attn.heads,
                                          attention_mask)
