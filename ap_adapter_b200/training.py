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


class BucketedGradReducer:
    """Data-parallel gradient averaging of the adapter weights, overlapped with the backward pass (what torch DDP does for the
    reference run under accelerate, train_apadapter_v2.py:939-958): the 64 adapter gradients are grouped, in the order the backward
    produces them (last site first), into buckets of ``bucket_bytes``; a post-accumulate-grad hook copies each gradient into its
    bucket and, when a bucket is complete, issues ONE asynchronous ``all_reduce`` for it — NCCL runs it on its own stream over
    NVLink while autograd is still computing the earlier sites.  ``finish()`` waits, divides by the world size and writes the
    averages back into ``p.grad``.  Outside a process group (or with one rank) it does nothing."""

    def __init__(self, params: Iterable[nn.Parameter], bucket_bytes: int = 25 << 20, group=None):
        import torch.distributed as dist

        self.group = group
        self.params = list(params)
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.buckets = []            # (flat fp32 buffer, [(param, offset, numel)])
        self.slot = {}               # id(param) -> (bucket index, offset)
        self.handles, self.pending, self.hooks = [], [], []
        self.bytes_reduced = 0
        if self.world == 1 or not self.params:
            return
        cur, cur_n = [], 0
        for p in reversed(self.params):                       # backward order
            if cur and (cur_n + p.numel()) * 4 > bucket_bytes:
                self._close(cur, cur_n)
                cur, cur_n = [], 0
            cur.append((p, cur_n, p.numel()))
            cur_n += p.numel()
        if cur:
            self._close(cur, cur_n)
        for p in self.params:
            self.hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _close(self, entries, n):
        buf = torch.zeros(n, dtype=torch.float32, device=entries[0][0].device)
        for p, off, _ in entries:
            self.slot[id(p)] = (len(self.buckets), off)
        self.buckets.append((buf, entries))
        self.pending.append(len(entries))

    def _on_grad(self, p):
        import torch.distributed as dist

        bi, off = self.slot[id(p)]
        buf, entries = self.buckets[bi]
        buf[off:off + p.numel()].copy_(p.grad.reshape(-1))
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self.handles.append(dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self.bytes_reduced += buf.numel() * 4

    def finish(self) -> int:
        """Wait for the bucket all-reduces, average, write back.  Returns the bytes reduced in this step."""
        if self.world == 1:
            return 0
        for h in self.handles:
            h.wait()
        for bi, (buf, entries) in enumerate(self.buckets):
            if self.pending[bi] != 0:
                continue            # a parameter of this bucket got no gradient this step: nothing was reduced for it
            buf.div_(self.world)
            for p, off, n in entries:
                p.grad.copy_(buf[off:off + n].view_as(p.grad))
        done = self.bytes_reduced
        self.handles, self.bytes_reduced = [], 0
        self.pending = [len(e) for _, e in self.buckets]
        return done

    def remove(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []


def train_step(unet: nn.Module, scheduler: DDIMScheduler, optimizer: torch.optim.Optimizer, latents: torch.Tensor,
               generated_prompt_embeds: torch.Tensor, prompt_embeds: torch.Tensor, attention_mask: Optional[torch.Tensor],
               max_grad_norm: float = 1.0, autocast_dtype: Optional[torch.dtype] = torch.bfloat16, generator=None,
               params: Optional[List[nn.Parameter]] = None, group=None, reducer: Optional[BucketedGradReducer] = None) -> float:
    """One optimiser step (train_apadapter_v2.py:895-979).  ``latents`` (B, 8, T/4, 16) stand for ``vae.encode(mel) * scaling``;
    ``generated_prompt_embeds`` (B, 8 + La, 768) already carries the AudioMAE tokens behind the 8 text tokens.  ``reducer`` (a
    ``BucketedGradReducer`` over ``params``) overlaps the gradient all-reduce with the backward; without it the gradients are
    reduced in one flat all-reduce after the backward.  Returns the loss."""
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
    if reducer is not None:
        reducer.finish()          # the bucket all-reduces were issued from autograd hooks during backward()
    else:
        allreduce_adapter_grads(params, group)
    if max_grad_norm is not None:
        torch.nn.utils.clip_grad_norm_(params, max_grad_norm)
    optimizer.step()
    return float(loss.detach())
