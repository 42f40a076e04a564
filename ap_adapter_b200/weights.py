"""On-disk weight formats of the reference, read as they are written there.

* ``copied_cross_attention/<site>_{k,v}.bin`` — one pickled fp16 ``nn.Parameter`` (C, 768) per file, written by the
  reference's ``copy_weight.py:61-62`` and read by ``IPAttnProcessor2_0(do_copy=True)``
  (``attention_processor.py:328-344``).  ``<site>`` is the processor key, e.g.
  ``down_blocks.1.attentions.1.transformer_blocks.0.attn2.processor``.
* adapter checkpoint ``pytorch_model.bin`` — a flat state dict with keys ``<site>.to_{k,v}_ip.weight``
  (``inference.py:51-57``, ``train_apadapter_v2.py:656-662``), as saved by ``accelerator.save_state`` from the
  ``AttnProcsLayers`` wrapper.
* AudioMAE ``pretrained.pth`` — ``{'model': state_dict}`` with timm key names (``AudioMAE.py:28-30``).
"""
from __future__ import annotations

import os
from typing import Dict, Mapping, Tuple

import torch
import torch.nn as nn


def load_copied_cross_attention(directory: str, name: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (k_weight, v_weight) of one adapter site.  The files were pickled from CUDA tensors, so they are
    mapped to the CPU on load (the reference's bare ``torch.load`` needs a GPU)."""
    out = []
    for suffix in ("k", "v"):
        path = os.path.join(directory, f"{name}_{suffix}.bin")
        w = torch.load(path, map_location="cpu", weights_only=False)
        out.append(w.detach() if isinstance(w, torch.Tensor) else torch.as_tensor(w))
    return out[0], out[1]


def save_copied_cross_attention(directory: str, name: str, k_weight: torch.Tensor, v_weight: torch.Tensor) -> None:
    """Writer for the same format (what ``copy_weight.py`` produces from ``attn2.to_k`` / ``attn2.to_v``)."""
    os.makedirs(directory, exist_ok=True)
    torch.save(nn.Parameter(k_weight.detach().clone(), requires_grad=False), os.path.join(directory, f"{name}_k.bin"))
    torch.save(nn.Parameter(v_weight.detach().clone(), requires_grad=False), os.path.join(directory, f"{name}_v.bin"))


def adapter_state_dict(attn_procs: Mapping[str, nn.Module]) -> Dict[str, torch.Tensor]:
    """``{<site>.to_k_ip.weight, <site>.to_v_ip.weight}`` for every adapter processor in ``attn_procs``."""
    sd = {}
    for name, proc in attn_procs.items():
        if hasattr(proc, "to_k_ip") and hasattr(proc, "to_v_ip"):
            sd[name + ".to_k_ip.weight"] = proc.to_k_ip.weight.detach()
            sd[name + ".to_v_ip.weight"] = proc.to_v_ip.weight.detach()
    return sd


def load_adapter_state_dict(attn_procs: Mapping[str, nn.Module], state_dict: Mapping[str, torch.Tensor],
                            dtype: torch.dtype | None = torch.float16) -> int:
    """What ``inference.py:51-57`` does: assign ``state_dict[name + '.to_{k,v}_ip.weight']`` to every adapter
    processor (as new Parameters, cast to ``dtype``).  Returns the number of sites loaded; a missing key raises."""
    n = 0
    for name, proc in attn_procs.items():
        if hasattr(proc, "to_v_ip") or hasattr(proc, "to_k_ip"):
            wk = state_dict[name + ".to_k_ip.weight"]
            wv = state_dict[name + ".to_v_ip.weight"]
            dev = proc.to_k_ip.weight.device
            proc.to_v_ip.weight = nn.Parameter(wv.to(device=dev, dtype=dtype or wv.dtype))
            proc.to_k_ip.weight = nn.Parameter(wk.to(device=dev, dtype=dtype or wk.dtype))
            n += 1
    return n


def load_adapter_checkpoint(attn_procs: Mapping[str, nn.Module], path: str, dtype: torch.dtype | None = torch.float16) -> int:
    return load_adapter_state_dict(attn_procs, torch.load(path, map_location="cpu", weights_only=False), dtype)


def load_audiomae_checkpoint(path: str) -> Dict[str, torch.Tensor]:
    """``torch.load(path)['model']`` — encoder keys are used, decoder / mask-token keys are ignored
    (the reference loads with ``strict=False``, AudioMAE.py:28-30)."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    sd = ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt else ckpt
    return {k: v for k, v in sd.items() if not k.startswith(("decoder", "mask_token"))}
