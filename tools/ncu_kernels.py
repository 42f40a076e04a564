"""Drive every libapx kernel once at a representative size so that ONE ncu run captures them all (test infrastructure).
    ncu --set full --clock-control none -k regex:'gemm_tc|ipattn|groupnorm|...' -o /tmp/prof_all python tools/ncu_kernels.py
    ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_all_raw.csv      (the report itself exceeds gpurun's 64 MiB return limit)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ap_adapter_b200 import _lib  # noqa: E402
from ap_adapter_b200.audiomae import AudioMAEConditionCTPoolRand, extract_kaldi_fbank_feature_batch  # noqa: E402
from ap_adapter_b200.unet import FeedForward  # noqa: E402
from gpu_util import build_site, random_site_case  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
lib = _lib.load()
# audio front-end: polyphase resample (44.1 kHz -> 16 kHz, 160 phases) + batched Kaldi fbank, 4 clips of 10 s
wav = torch.randn(4, 441000, device=dev) * 0.1
extract_kaldi_fbank_feature_batch(wav, 44100, 1024)
# AudioMAE encode + pool, 8 clips (patchify, gemm16 / gemm_tc incl. folded LayerNorm, flash_attn, layernorm, pool_tokens)
os.environ["APX_MAE_GRAPH"] = "0"
from ap_adapter_b200.audiomae import MaskedAutoencoderViTEncoder  # noqa: E402
MaskedAutoencoderViTEncoder.use_graph = False
enc = AudioMAEConditionCTPoolRand(checkpoint_path=None, depth=1).to(dev).half().eval()
mel = torch.randn(8, 1024, 128, device=dev) * 0.45 - 0.45
with torch.no_grad():
    enc(mel, time_pool=2, freq_pool=2)
# UNet-block neighbours: GroupNorm(+bias+temb)+SiLU at the largest and a mid-size group, the ResnetBlock output add, the FeedForward
# with norm3 folded into the GEGLU GEMM (S = 1000, C = 256)
cl = lambda t: t.contiguous(memory_format=torch.channels_last)  # noqa: E731
for (b, c, h, w) in ((2, 256, 250, 16), (2, 384, 63, 4), (16, 256, 125, 8)):
    x = cl(torch.randn(b, c, h, w, device=dev, dtype=torch.float16))
    g = torch.ones(c, device=dev)
    _lib.groupnorm_nhwc(x, g, g, 32, 1e-5, True, g, torch.randn(b, c, device=dev, dtype=torch.float16))
    _lib.add_bias_nhwc(x, x, g, 1.0)
ff = FeedForward(256).to(dev).half().eval()
norm = torch.nn.LayerNorm(256).to(dev).half()
with torch.no_grad():
    ff.fused_ln(torch.randn(2, 1000, 256, device=dev, dtype=torch.float16), norm)
# adapter sites at the three real geometries: Bc = 2 (all-heads cluster launch / one-head cluster kernel / three kernels) and
# Bc = 16 (all-heads kernel, one tile per SM), forward + backward
for (bc, s, c, la) in ((2, 1000, 256, 128), (2, 252, 384, 128), (2, 64, 640, 128), (16, 1000, 256, 128)):
    case = random_site_case(bc, s, c, 8, 8, la, scale=1.0, seed=1)
    attn, proc = build_site(case, torch.float16)
    for p in attn.parameters():
        p.requires_grad_(False)
    x = case["x"].to(dev, torch.float16).requires_grad_(bc == 16)
    y = proc(attn, x, encoder_hidden_states=case["ctx"].to(dev, torch.float16))
    if bc == 16:
        y.backward(torch.ones_like(y))
# the opt-in variants: one-head kernel at C = 640, four-warpgroup all-heads kernel at Bc = 16
for variant, (bc, s, c) in ((_lib.APX_SITE_ONEHEAD, (2, 64, 640)), (_lib.APX_SITE_ALLHEADS4, (16, 1000, 256))):
    lib.apx_set_site_variant(variant)
    case = random_site_case(bc, s, c, 8, 8, 128, scale=1.0, seed=1)
    attn, proc = build_site(case, torch.float16)
    with torch.no_grad():
        proc(attn, case["x"].to(dev, torch.float16), encoder_hidden_states=case["ctx"].to(dev, torch.float16))
    lib.apx_set_site_variant(_lib.APX_SITE_AUTO)
# flash-style self-attention: the UNet's S = 1000 sites (d = 32, one wave / batch 8) and the ViT's 513 keys (d = 64)
for (b, s, c, h) in ((2, 1000, 256, 8), (16, 1000, 256, 8), (8, 513, 768, 12)):
    qkv = torch.randn(b, s, 3 * c, device=dev, dtype=torch.float16)
    q, k, v = qkv.chunk(3, dim=-1)
    _lib.flash_attention(q, k, v, h)
# the token GEMMs of one denoise step through apx_gemm16 (direct TMA-store epilogue / cluster split-K finisher)
for (m, n, k, kind) in ((2000, 768, 256, "ln"), (504, 384, 1536, "res"), (128, 640, 2560, "res"), (128, 5120, 640, "ln_geglu"), (2000, 256, 256, "res")):
    a = torch.randn(m, k, device=dev, dtype=torch.float16)
    w = torch.randn(n, k, device=dev, dtype=torch.float16) * 0.05
    b16 = torch.randn(n, device=dev, dtype=torch.float16)
    if kind.startswith("ln"):
        wg, c1, c2 = _lib.fold_layernorm(w, b16, torch.ones(k, device=dev, dtype=torch.float16), torch.zeros(k, device=dev, dtype=torch.float16))
        _lib.linear_lnfold(a, wg, c1, c2, 1e-5, geglu=kind.endswith("geglu"))
    else:
        _lib.linear(a, w, b16.float(), residual=torch.randn(m, n, device=dev, dtype=torch.float16))
torch.cuda.synchronize()
print("done")
