"""Adapter-only training step (BASELINE config 5) around the sm_100a processors.

Restates the optimisation step of the reference's ``train_apadapter_v2.py`` (``:892-979``): sample noise and per-clip
timesteps, ``add_noise``, UNet forward with the adapter processors, MSE against epsilon (or v), backward, gradient-norm clip,
AdamW step — with ONLY ``to_k_ip`` / ``to_v_ip`` of the 32 adapter sites trainable (``:612-669, 759-769``).  The reference
delegates data parallelism to HF accelerate (torch DDP over NCCL); here it is explicit and minimal: every rank runs its own
clips, and the 64 adapter gradients (21.6 M fp32 values, 86.5 MB) are averaged with ONE flat ``all_reduce`` per optimiser step
(``allreduce_adapter_grads``).  VAE encoding, prompt encoding and the dataset are out of scope: latents and
the conditioning tensors are inputs.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .pipeline import DDIMScheduler


def adapter_parameters(unet: nn.Module) -> List[nn.Parameter]:
    """The trainable tensors of the reference run: ``to_k_ip.weight`` / ``to_v_ip.weight`` of every adapter processor, in
    registry order (train_apadapter_v2.py:665-669 wraps exactly these in ``AttnProcsLayers``)."""
    params: List[nn.Parameter] = []
    for _, proc in unet.attn_processors.items():
        if hasattr(proc, "to_k_ip"):
            params += [proc.to_k_ip.weight, proc.to_v_ip.weight]
    return params


def freeze_all_but_adapter(unet: nn.Module) -> List[nn.Parameter]:
    """``unet.requires_grad_(False)`` then re-enable the adapter weights (train_apadapter_v2.py:590-612, 646-648)."""
    for p in unet.parameters():
        p.requires_grad_(False)
    params = adapter_parameters(unet)
    for p in params:
        p.requires_grad_(True)
    return params


def add_noise(scheduler: DDIMScheduler, latents: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
    """HF ``SchedulerMixin.add_noise``: ``sqrt(a_t) x0 + sqrt(1 - a_t) eps`` with per-sample timesteps."""
    a = scheduler.alphas_cumprod.to(device=latents.device, dtype=torch.float32)[timesteps]
    sa = a.sqrt().view(-1, 1, 1, 1)
    s1a = (1.0 - a).sqrt().view(-1, 1, 1, 1)
    return (sa * latents.float() + s1a * noise.float()).to(latents.dtype)


def get_velocity(scheduler: DDIMScheduler, latents: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
    a = scheduler.alphas_cumprod.to(device=latents.device, dtype=torch.float32)[timesteps]
    sa = a.sqrt().view(-1, 1, 1, 1)
    s1a = (1.0 - a).sqrt().view(-1, 1, 1, 1)
    return (sa * noise.float() - s1a * latents.float()).to(latents.dtype)


def allreduce_adapter_grads(params: Iterable[nn.Parameter], group=None) -> int:
    """Average the adapter gradients over the ranks with one flat all-reduce (the DDP bucket of the reference run, 86.5 MB in
    fp32).  Returns the number of bytes reduced.  No-op outside a process group."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel() * 4


def train_step(unet: nn.Module, scheduler: DDIMScheduler, optimizer: torch.optim.Optimizer, latents: torch.Tensor,
               generated_prompt_embeds: torch.Tensor, prompt_embeds: torch.Tensor, attention_mask: Optional[torch.Tensor],
               max_grad_norm: float = 1.0, autocast_dtype: Optional[torch.dtype] = torch.bfloat16, generator=None,
               params: Optional[List[nn.Parameter]] = None, group=None) -> float:
    """One optimiser step (train_apadapter_v2.py:895-979).  ``latents`` (B, 8, T/4, 16) stand for ``vae.encode(mel) * scaling``;
    ``generated_prompt_embeds`` (B, 8 + La, 768) already carries the AudioMAE tokens behind the 8 text tokens.  Returns the loss."""
    params = adapter_parameters(unet) if params is None else params
    noise = torch.randn(latents.shape, generator=generator, device=latents.device, dtype=latents.dtype)
    timesteps = torch.randint(0, scheduler.num_train_timesteps, (latents.shape[0],), generator=generator,
                              device=latents.device).long()
    noisy = add_noise(scheduler, latents, noise, timesteps)
    with torch.autocast(latents.device.type, dtype=autocast_dtype, enabled=autocast_dtype is not None):
        pred = unet(noisy, timesteps, encoder_hidden_states=generated_prompt_embeds, encoder_hidden_states_1=prompt_embeds,
                    encoder_attention_mask_1=attention_mask, return_dict=False)[0]
    if scheduler.prediction_type == "epsilon":
        target = noise
    elif scheduler.prediction_type == "v_prediction":
        target = get_velocity(scheduler, latents, noise, timesteps)
    else:
        raise ValueError(f"Unknown prediction type {scheduler.prediction_type}")
    loss = F.mse_loss(pred.float(), target.float(), reduction="mean")
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    allreduce_adapter_grads(params, group)
    if max_grad_norm is not None:
        torch.nn.utils.clip_grad_norm_(params, max_grad_norm)
    optimizer.step()
    return float(loss.detach())
