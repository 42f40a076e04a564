"""On a bad run of the LN + GEGLU 64-column-tile GEMM: which perturbed row statistics explain the observed outputs?"""
import itertools
import torch, torch.nn.functional as F
from ap_adapter_b200 import _lib

m, k, n = 504, 384, 3072
g = torch.Generator().manual_seed(m + n)
x = (torch.randn(m, k, generator=g) * 1.5 + 0.8 * torch.randn(m, 1, generator=g) + 0.3).half()
w = (torch.randn(n, k, generator=g) * k ** -0.5).half()
b = torch.randn(n, generator=g) * 0.2
gamma, beta = 1.0 + 0.3 * torch.randn(k, generator=g), 0.2 * torch.randn(k, generator=g)
pre = F.linear(F.layer_norm(x.float(), (k,), gamma, beta, 1e-5), w.float(), b)
ref = pre[:, 0::2] * F.gelu(pre[:, 1::2])
xc, wc, bc = x.cuda(), w.cuda(), b.cuda()
wg, c1, c2 = _lib.fold_layernorm(wc, bc, gamma.cuda(), beta.cuda())
wgf, c1f, c2f = wg.float().cpu(), c1.cpu(), c2.cpu()
junk = torch.empty(64 << 20, device="cuda", dtype=torch.uint8)
nkb = k // 64


def predict(row, cols, mean, var):
    acc = x[row].float() @ wgf[cols].t()
    p = (acc - mean * c1f[cols]) * torch.rsqrt(var + 1e-5) + c2f[cols]
    return p[0::2] * F.gelu(p[1::2])


def half_stats(row, hf, kbs):
    """(mean, M2, n) of the elements thread (row, hf) reads when it sees k-blocks `kbs` (in that order)"""
    xs = torch.cat([x[row, kb * 64 + hf * 32: kb * 64 + hf * 32 + 32].float() for kb in kbs])
    return xs.mean(), ((xs - xs.mean()) ** 2).sum(), xs.numel()


found = 0
for r in range(400):
    if r % 3 == 0:
        junk.random_(0, 255)
    y = _lib.linear_lnfold(x.cuda() if r % 2 == 0 else xc, wg, c1, c2, 1e-5, geglu=True).float().cpu()
    e = (y - ref).abs()
    d = (e > 2e-3 * ref.abs().max()).nonzero()
    if not len(d):
        continue
    row, col = d[0].tolist()
    tile = col // 32
    cols = torch.arange(tile * 64, tile * 64 + 64)
    obs = y[row, tile * 32: tile * 32 + 32]
    exact = predict(row, cols, x[row].float().mean(), x[row].float().var(unbiased=False))
    print(f"run {r}: row {row} (tile row {row % 128}) col tile {tile}; err vs exact stats {((obs - exact).abs().max()).item():.3e}")
    best = []
    seqs = [list(range(nkb))]
    for i in range(nkb):
        for j in range(nkb):
            if i != j:
                s = list(range(nkb)); s[i] = j; seqs.append(s)
    for hf in (0, 1):
        for s in seqs:
            ma, M2a, na = half_stats(row, hf, s)
            mb, M2b, nb = half_stats(row, 1 - hf, list(range(nkb)))
            mean = 0.5 * (ma + mb)
            var = (M2a + M2b + (mb - ma) ** 2 * na / 2) / k
            err = (obs - predict(row, cols, mean, var)).abs().max().item()
            best.append((err, hf, s))
    best.sort(key=lambda t: t[0])
    for err, hf, s in best[:3]:
        print(f"    half {hf} saw k-blocks {s}: max abs diff to the observed outputs {err:.3e}")
    found += 1
    if found >= 6:
        break
print("bad runs analysed:", found)
