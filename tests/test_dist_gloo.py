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
