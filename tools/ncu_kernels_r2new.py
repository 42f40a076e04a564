"""The kernels written in the second half of round 2, once each at the sizes the denoise step / the ViT run them, for ONE ncu run
(test infrastructure; the full driver is tools/ncu_kernels.py):
    ncu --section SpeedOfLight --section LaunchStats --section Occupancy --section WarpStateStats --section MemoryWorkloadAnalysis \
        --section ComputeWorkloadAnalysis --clock-control none \
        -k regex:'flash_attn|gemm16_kernel|groupnorm_nhwc_v2|add_layernorm16|bwd_tc|wgrad' -o /tmp/prof python tools/ncu_kernels_r2new.py
Every section is guarded: a failure in one does not lose the captures of the others.
"""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ap_adapter_b200 import _lib  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
lib = _lib.load()
f16 = torch.float16


def section(fn):
    try:
        fn()
        torch.cuda.synchronize()
        print("ok  ", fn.__name__, flush=True)
    except Exception:   # noqa: BLE001
        print("FAIL", fn.__name__, flush=True)
        traceback.print_exc()


def flash():
    # the UNet's S = 1000 self-attention sites (d = 32; B = 2: one wave, B = 16: batch 8 for reference) and the ViT's 513 keys (d = 64)
    for (b, s, c, h) in ((2, 1000, 256, 8), (8, 513, 768, 12), (16, 1000, 256, 8)):
        qkv = torch.randn(b, s, 3 * c, device=dev, dtype=f16)
        q, k, v = qkv.chunk(3, dim=-1)
        _lib.flash_attention(q, k, v, h)


def gemm16():
    # the token GEMMs of one denoise step (direct TMA-store epilogue / cluster split-K finisher) and AudioMAE's fc1 at 8 clips
    for (m, n, k, kind) in ((2000, 2048, 256, "ln_geglu"), (2000, 768, 256, "ln"), (2000, 256, 1024, "res"), (504, 384, 1536, "res"),
                            (128, 640, 2560, "res"), (128, 5120, 640, "ln_geglu"), (4104, 3072, 768, "ln_gelu")):
        a = torch.randn(m, k, device=dev, dtype=f16)
        w = torch.randn(n, k, device=dev, dtype=f16) * 0.05
        b16 = torch.randn(n, device=dev, dtype=f16)
        if kind.startswith("ln"):
            wg, c1, c2 = _lib.fold_layernorm(w, b16, torch.ones(k, device=dev, dtype=f16), torch.zeros(k, device=dev, dtype=f16))
            _lib.linear_lnfold(a, wg, c1, c2, 1e-5, geglu=kind.endswith("geglu"), gelu=kind.endswith("gelu"))
        else:
            _lib.linear(a, w, b16.float(), residual=torch.randn(m, n, device=dev, dtype=f16))


def groupnorm():
    # GroupNorm(+conv bias + time embedding)+SiLU, second form: the largest group of the step, a mid-size one, batch 8
    for (b, c, h, w) in ((2, 256, 250, 16), (2, 384, 63, 4), (2, 640, 32, 1), (16, 256, 125, 8)):
        x = torch.randn(b, c, h, w, device=dev, dtype=f16).contiguous(memory_format=torch.channels_last)
        g = torch.ones(c, device=dev)
        _lib.groupnorm_nhwc(x, g, g, 32, 1e-5, True, g, torch.randn(b, c, device=dev, dtype=f16))


def add_ln():
    # residual add + LayerNorm (and the LayerNorm-only form) at the sites below 1000 rows
    for (rows, c) in ((504, 384), (128, 640)):
        norm = torch.nn.LayerNorm(c).to(dev).half()
        a = torch.randn(2, rows // 2, c, device=dev, dtype=f16)
        _lib.add_layernorm(a, torch.randn_like(a), norm)
        _lib.add_layernorm(a, None, norm)


def backward():
    # adapter-site backward at batch 8 (Bc = 16, S = 1000, C = 256, La = 128): tcgen05 attention backward + weight gradients
    from gpu_util import build_site, random_site_case
    case = random_site_case(16, 1000, 256, 8, 8, 128, scale=1.0, seed=1)
    attn, proc = build_site(case, f16)
    for p in attn.parameters():
        p.requires_grad_(False)
    x = case["x"].to(dev, f16).requires_grad_(True)
    y = proc(attn, x, encoder_hidden_states=case["ctx"].to(dev, f16))
    y.backward(torch.ones_like(y))


for fn in (flash, gemm16, groupnorm, add_ln, backward):
    section(fn)
print("done", flush=True)
