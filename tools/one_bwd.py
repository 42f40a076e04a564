import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from gpu_util import build_site, random_site_case
case = random_site_case(16, 1000, 256, 8, 8, 128, scale=1.0, seed=1)
attn, proc = build_site(case, torch.bfloat16)
for p in attn.parameters(): p.requires_grad_(False)
x = case["x"].to("cuda", torch.bfloat16).requires_grad_(True)
y = proc(attn, x, encoder_hidden_states=case["ctx"].to("cuda", torch.bfloat16)); y.backward(torch.randn_like(x)); torch.cuda.synchronize()
