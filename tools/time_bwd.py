"""Device time of the adapter backward (apx_ipattn_bwd) per site geometry, CUDA-event timed (test infrastructure)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_util import build_site, random_site_case  # noqa: E402

for (bc, s, c, la) in ((16, 1000, 256, 128), (16, 252, 384, 128), (16, 64, 640, 128), (2, 1000, 256, 128)):
    case = random_site_case(bc, s, c, 8, 8, la, scale=1.0, seed=1)
    attn, proc = build_site(case, torch.bfloat16)
    for p in attn.parameters():
        p.requires_grad_(False)
    x = case["x"].to("cuda", torch.bfloat16).requires_grad_(True)
    enc = case["ctx"].to("cuda", torch.bfloat16)
    dy = torch.randn_like(x)
    for _ in range(2):
        y = proc(attn, x, encoder_hidden_states=enc)
        y.backward(dy)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    n = 10
    tf = tb = 0.0
    for _ in range(n):
        e[0].record()
        y = proc(attn, x, encoder_hidden_states=enc)
        e[1].record()
        y.backward(dy)
        e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    print(f"Bc={bc} S={s} C={c} La={la}: forward (ctx cache + site) {1e3 * tf / n:.1f} us, backward {1e3 * tb / n:.1f} us")
