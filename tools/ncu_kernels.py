"""Drive every libapx kernel once at a representative size so that ONE ncu run captures them all (test infrastructure).
    ncu --set full --clock-control none -k regex:'apx::' -o gpurun_out/prof_all python tools/ncu_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from ap_adapter_b200.audiomae import AudioMAEConditionCTPoolRand  # noqa: E402
from gpu_util import build_site, random_site_case  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
# AudioMAE encode + pool, 8 clips (patch_embed, layernorm, gemm_tc<64/128>, kv_pack_views, ipattn_core, pool_tokens)
enc = AudioMAEConditionCTPoolRand(checkpoint_path=None, depth=2).to(dev).half().eval()
mel = torch.randn(8, 1024, 128, device=dev) * 0.45 - 0.45
with torch.no_grad():
    enc(mel, time_pool=2, freq_pool=2)
# adapter sites at the three real geometries, Bc = 16 (site kernel, three-kernel path), forward + backward
for (bc, s, c, la) in ((16, 1000, 256, 128), (16, 252, 384, 128), (16, 64, 640, 128)):
    case = random_site_case(bc, s, c, 8, 8, la, scale=1.0, seed=1)
    attn, proc = build_site(case, torch.float16)
    for p in attn.parameters():
        p.requires_grad_(False)
    x = case["x"].to(dev, torch.float16).requires_grad_(True)
    y = proc(attn, x, encoder_hidden_states=case["ctx"].to(dev, torch.float16))
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("done")
