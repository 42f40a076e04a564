"""Batch-sharded inference across ranks: world_size 2 on the gloo backend, CPU.

Each rank denoises its contiguous shard of clips with the same weights; the only collective is the gather of final
latents.  The gathered result must equal the single-process result clip for clip."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ap_adapter_b200.pipeline import APAdapterDenoiser, gather_latents, shard_range
from ap_adapter_b200.unet import AudioLDM2UNet2DConditionModel, build_attn_processors
from oracle.ref_processors import OracleAttnProcessor2_0, OracleIPAttnProcessor2_0

TINY = dict(block_out_channels=(32, 64, 96, 128), norm_num_groups=16)
N_CLIPS, STEPS = 3, 2


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    torch.manual_seed(0)
    unet = AudioLDM2UNet2DConditionModel(**TINY).eval()
    unet.set_attn_processor(build_attn_processors(unet, OracleIPAttnProcessor2_0, OracleAttnProcessor2_0, ap_scale=0.5))
    g = torch.Generator().manual_seed(1)
    lat = torch.randn(N_CLIPS, 8, 24, 16, generator=g)
    gpe = torch.randn(2 * N_CLIPS, 8 + 8, 768, generator=g)   # rows [0,N) uncond, [N,2N) cond
    pe = torch.randn(2 * N_CLIPS, 4, 1024, generator=g)
    mask = torch.ones(2 * N_CLIPS, 4)
    return unet, lat, gpe, pe, mask


def _shard(t2n, lo, hi):
    """rows of a [uncond; cond] (2N, ...) tensor that belong to clips [lo, hi)."""
    n = t2n.shape[0] // 2
    return torch.cat([t2n[lo:hi], t2n[n + lo:n + hi]], dim=0)


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        unet, lat, gpe, pe, mask = _problem()
        lo, hi = shard_range(N_CLIPS, rank, world)
        den = APAdapterDenoiser(unet)
        local = den.denoise(lat[lo:hi], _shard(gpe, lo, hi), _shard(pe, lo, hi), _shard(mask, lo, hi),
                            num_inference_steps=50, guidance_scale=3.5, use_cuda_graph=False, steps_to_run=STEPS)
        full = gather_latents(local, N_CLIPS)
        if rank == 0:
            torch.save(full, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_denoise_equals_single_process(tmp_path):
    out = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    gathered = torch.load(out)
    unet, lat, gpe, pe, mask = _problem()
    single = APAdapterDenoiser(unet).denoise(lat, gpe, pe, mask, num_inference_steps=50, guidance_scale=3.5,
                                             use_cuda_graph=False, steps_to_run=STEPS)
    assert gathered.shape == single.shape == (N_CLIPS, 8, 24, 16)
    # per-clip arithmetic is independent of the batch it runs in (up to GEMM blocking on the CPU)
    assert ((gathered - single).abs().max() / single.abs().max()).item() < 1e-5


# ---- adapter training, data parallel (BASELINE config 5): one flat all-reduce of the 64 adapter gradients per step -----------
def _train_worker(rank, world, port, out_path):
    from ap_adapter_b200 import training

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        unet, lat, gpe, pe, mask = _problem()          # identical initial weights on every rank (same seed)
        unet.train(False)
        params = training.freeze_all_but_adapter(unet)
        assert len(params) == 64
        w0 = [p.detach().clone() for p in params]
        opt = torch.optim.SGD(params, lr=1e-1)
        den = APAdapterDenoiser(unet)
        lo, hi = shard_range(2, rank, world)            # one clip per rank
        g = torch.Generator().manual_seed(100 + rank)   # different noise / timesteps per rank
        # local gradients first (no collective), to know what the average must be
        loss = training.train_step(unet, den.scheduler, torch.optim.SGD(params, lr=0.0), lat[lo:hi], gpe[lo:hi], pe[lo:hi],
                                   mask[lo:hi], autocast_dtype=None, generator=g, params=params, max_grad_norm=None,
                                   group=dist.new_group([rank]))
        local = torch.cat([p.grad.reshape(-1) for p in params])
        both = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(both, local)
        want = sum(both) / world
        # the real step: same noise (generator re-seeded), gradients averaged over the ranks, then SGD
        g.manual_seed(100 + rank)
        loss2 = training.train_step(unet, den.scheduler, opt, lat[lo:hi], gpe[lo:hi], pe[lo:hi], mask[lo:hi], autocast_dtype=None,
                                    generator=g, params=params, max_grad_norm=None)
        got = torch.cat([p.grad.reshape(-1) for p in params])
        assert abs(loss - loss2) < 1e-6 * max(1.0, abs(loss))
        assert ((got - want).abs().max() / want.abs().max().clamp_min(1e-30)).item() < 1e-5
        assert any((p.detach() - w).abs().max() > 0 for p, w in zip(params, w0))       # the step moved the adapter
        torch.save(torch.cat([p.detach().reshape(-1) for p in params]), f"{out_path}.{rank}")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_adapter_training_averages_gradients(tmp_path):
    out = str(tmp_path / "weights.pt")
    mp.spawn(_train_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    w0, w1 = torch.load(out + ".0"), torch.load(out + ".1")
    assert torch.equal(w0, w1)   # replicas stay bit-identical after the averaged step


def _bucket_worker(rank, world, port, out_path):
    from ap_adapter_b200 import training

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        unet, lat, gpe, pe, mask = _problem()
        params = training.freeze_all_but_adapter(unet)
        den = APAdapterDenoiser(unet)
        lo, hi = shard_range(2, rank, world)
        g = torch.Generator().manual_seed(100 + rank)
        opt = torch.optim.SGD(params, lr=0.0)
        # reference: one flat all-reduce after the backward
        training.train_step(unet, den.scheduler, opt, lat[lo:hi], gpe[lo:hi], pe[lo:hi], mask[lo:hi], autocast_dtype=None, generator=g,
                            params=params, max_grad_norm=None)
        want = torch.cat([p.grad.reshape(-1) for p in params])
        # bucketed: tiny buckets so that the 64 gradients spread over many all-reduces issued from the autograd hooks
        red = training.BucketedGradReducer(params, bucket_bytes=64 << 10)
        assert len(red.buckets) > 4 and sum(len(e) for _, e in red.buckets) == 64
        for _ in range(2):                       # twice: the reducer re-arms itself
            g.manual_seed(100 + rank)
            training.train_step(unet, den.scheduler, opt, lat[lo:hi], gpe[lo:hi], pe[lo:hi], mask[lo:hi], autocast_dtype=None,
                                generator=g, params=params, max_grad_norm=None, reducer=red)
            got = torch.cat([p.grad.reshape(-1) for p in params])
            assert ((got - want).abs().max() / want.abs().max().clamp_min(1e-30)).item() < 1e-6
        red.remove()
        if rank == 0:
            torch.save(got, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_bucketed_overlapped_all_reduce_equals_flat_all_reduce(tmp_path):
    """BucketedGradReducer (all-reduces issued per bucket from autograd hooks, during the backward) averages the adapter gradients
    exactly as the single flat all-reduce after the backward does."""
    out = str(tmp_path / "g.pt")
    mp.spawn(_bucket_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.isfinite(torch.load(out)).all()
