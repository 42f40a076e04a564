import torch, torch.nn.functional as F
from ap_adapter_b200 import _lib
def run(m,k,n,geglu,dtype=torch.float16):
    g = torch.Generator().manual_seed(m + n)
    x = (torch.randn(m, k, generator=g) * 1.5 + 0.8 * torch.randn(m, 1, generator=g) + 0.3).to(dtype)
    w = (torch.randn(n, k, generator=g) * k ** -0.5).to(dtype)
    b = torch.randn(n, generator=g) * 0.2
    gamma, beta = 1.0 + 0.3 * torch.randn(k, generator=g), 0.2 * torch.randn(k, generator=g)
    pre = F.linear(F.layer_norm(x.float(), (k,), gamma, beta, 1e-5), w.float(), b)
    ref = pre[:, 0::2] * F.gelu(pre[:, 1::2]) if geglu else pre
    wg, c1, c2 = _lib.fold_layernorm(w.cuda(), b.cuda(), gamma.cuda(), beta.cuda())
    ys = [_lib.linear_lnfold(x.cuda(), wg, c1, c2, 1e-5, geglu=geglu).float().cpu() for _ in range(3)]
    print('deterministic', (ys[0]==ys[1]).all().item(), (ys[0]==ys[2]).all().item())
    y = ys[0]
    e = (y-ref).abs()
    print(m,k,n,geglu, 'rel', (e.max()/ref.abs().max()).item())
    bad = (e > 2e-3*ref.abs().max()).nonzero()
    for r,c in bad.tolist()[:8]:
        print('  bad', r, c, 'got', y[r,c].item(), 'ref', ref[r,c].item(), 'hidden', pre[r,2*c].item() if geglu else None, 'gate', pre[r,2*c+1].item() if geglu else None)
    # the folded computation in fp32 torch, for comparison (same Wg rounding)
    xf = x.float(); mu = xf.mean(1, keepdim=True); var = xf.var(1, unbiased=False, keepdim=True)
    acc = xf @ wg.float().cpu().t()
    pre2 = (acc - mu * c1.cpu()[None]) * torch.rsqrt(var + 1e-5) + c2.cpu()[None]
    ref2 = pre2[:, 0::2] * F.gelu(pre2[:, 1::2]) if geglu else pre2
    e2 = (y - ref2).abs()
    print('  vs folded-fp32 emulation: rel', (e2.max()/ref.abs().max()).item(), ' emulation vs ref: rel', ((ref2-ref).abs().max()/ref.abs().max()).item())
run(504,384,3072,True)
import os
os.environ['APX_GEMM16']='0'
