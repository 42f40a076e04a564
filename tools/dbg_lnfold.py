import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from ap_adapter_b200 import _lib
for (m, k, n) in [(1026, 768, 2304), (1024, 768, 2304), (1026, 768, 256), (1026, 512, 2304), (2000, 256, 2048), (256, 768, 2304)]:
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(m, k, generator=g)).half().cuda()
    w = (torch.randn(n, k, generator=g) * k ** -0.5).half().cuda()
    b = torch.zeros(n).cuda(); gamma = torch.ones(k).cuda(); beta = torch.zeros(k).cuda()
    wg, c1, c2 = _lib.fold_layernorm(w, b, gamma, beta)
    y = _lib.linear_lnfold(x, wg, c1, c2, 1e-5)
    ref = F.linear(F.layer_norm(x.float(), (k,)), w.float())
    bad = ~torch.isfinite(y)
    rows = bad.any(1).nonzero().flatten()
    cols = bad.any(0).nonzero().flatten()
    err = ((y.float() - ref).abs().max() / ref.abs().max()).item() if not bad.any() else float('nan')
    print((m, k, n), "bad", int(bad.sum()), "rows", rows[:6].tolist(), len(rows), "cols", cols[:6].tolist(), len(cols), "err", err)
