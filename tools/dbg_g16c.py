import torch, torch.nn.functional as F
from ap_adapter_b200 import _lib
m,k,n=504,384,3072
g = torch.Generator().manual_seed(m + n)
x = (torch.randn(m, k, generator=g) * 1.5 + 0.8 * torch.randn(m, 1, generator=g) + 0.3).half()
w = (torch.randn(n, k, generator=g) * k ** -0.5).half()
b = torch.randn(n, generator=g) * 0.2
gamma, beta = 1.0 + 0.3 * torch.randn(k, generator=g), 0.2 * torch.randn(k, generator=g)
pre = F.linear(F.layer_norm(x.float(), (k,), gamma, beta, 1e-5), w.float(), b)
ref = pre[:, 0::2] * F.gelu(pre[:, 1::2])
for rep in range(4):
    wg, c1, c2 = _lib.fold_layernorm(w.cuda(), b.cuda(), gamma.cuda(), beta.cuda())
    ys = [_lib.linear_lnfold(x.cuda(), wg, c1, c2, 1e-5, geglu=True).float().cpu() for _ in range(3)]
    errs = [((y - ref).abs().max() / ref.abs().max()).item() for y in ys]
    print('rep', rep, 'errs', ' '.join(f'{e:.1e}' for e in errs))
    for y in ys:
        d = ((y - ref).abs() > 2e-3 * ref.abs().max()).nonzero()
        if len(d): print('   bad rows', sorted(set(d[:,0].tolist()))[:10], 'cols', min(d[:,1].tolist()), max(d[:,1].tolist()))
