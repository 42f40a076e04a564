"""Repeat-run stress of the 16-bit GEMM epilogues (apx_gemm16.cu): every run must sit within tolerance of fp32 torch, with
different preceding work / cache state between runs.  Run on the GPU box: PYTHONPATH=. python tools/stress_g16.py [runs]"""
import sys
import torch, torch.nn.functional as F
from ap_adapter_b200 import _lib


def case(m, k, n, ln, geglu, res, runs, dtype=torch.float16):
    g = torch.Generator().manual_seed(m + n)
    x = (torch.randn(m, k, generator=g) * 1.5 + 0.8 * torch.randn(m, 1, generator=g) + 0.3).to(dtype)
    w = (torch.randn(n, k, generator=g) * k ** -0.5).to(dtype)
    b = torch.randn(n, generator=g) * 0.2
    gamma, beta = 1.0 + 0.3 * torch.randn(k, generator=g), 0.2 * torch.randn(k, generator=g)
    n_out = n // 2 if geglu else n
    r = torch.randn(m, n_out, generator=g).to(dtype) if res else None
    xf = F.layer_norm(x.float(), (k,), gamma, beta, 1e-5) if ln else x.float()
    pre = F.linear(xf, w.float(), b)
    ref = pre[:, 0::2] * F.gelu(pre[:, 1::2]) if geglu else pre
    if res:
        ref = ref.to(dtype).float() + r.float()
    xc, wc, bc = x.cuda(), w.cuda(), b.cuda()
    rc = r.cuda() if res else None
    if ln:
        wg, c1, c2 = _lib.fold_layernorm(wc, bc, gamma.cuda(), beta.cuda())
    tol = (2e-3 if dtype == torch.float16 else 1.2e-2) * ref.abs().max()
    bad = 0; worst = 0.0; tiles = set(); first = None; differ = 0
    junk = torch.empty(64 << 20, device="cuda", dtype=torch.uint8)
    for it in range(runs):
        if it % 3 == 0:
            junk.random_(0, 255)
        xin = x.cuda() if it % 2 == 0 else xc
        y = (_lib.linear_lnfold(xin, wg, c1, c2, 1e-5, residual=rc, geglu=geglu) if ln else _lib.linear(xin, wc, bc, residual=rc, geglu=geglu)).float().cpu()
        if first is None:
            first = y
        differ += int(not torch.equal(first, y))
        e = (y - ref).abs()
        worst = max(worst, (e.max() / ref.abs().max()).item())
        d = (e > tol).nonzero()
        if len(d):
            bad += 1
            for rr, cc in d.tolist()[:50]:
                tiles.add((rr // 128, cc // 64))
    print(f"M={m} K={k} N={n} ln={int(ln)} geglu={int(geglu)} res={int(res)} {str(dtype)[6:]}: {bad}/{runs} bad, {differ} differ from run 0, worst rel err {worst:.2e} {sorted(tiles)[:8]}", flush=True)
    return bad + differ


if __name__ == "__main__":
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    total = 0
    for (m, k, n, ln, geglu, res) in [
            (2000, 256, 2048, True, True, False), (504, 384, 3072, True, True, False), (128, 640, 5120, True, True, False),   # FeedForward proj (norm3 folded)
            (2000, 1024, 256, False, False, True), (504, 1536, 384, False, False, True), (128, 2560, 640, False, False, True),  # FeedForward out + residual
            (2000, 256, 768, True, False, False), (504, 384, 1152, True, False, False), (128, 640, 1920, True, False, False),  # QKV (norm1 folded)
            (2000, 256, 256, False, False, True), (504, 384, 384, False, False, True), (128, 640, 640, False, False, True),    # to_out + residual
            (504, 384, 3072, False, True, False), (504, 384, 3072, True, False, False), (513, 768, 3072, True, False, False), (4104, 768, 2304, True, False, False)]:
        total += case(m, k, n, ln, geglu, res, runs)
    total += case(504, 384, 3072, True, True, False, runs, torch.bfloat16)
    total += case(128, 640, 1920, True, False, False, runs, torch.bfloat16)
    print("stress: FAILED" if total else "stress: ok")
    sys.exit(1 if total else 0)
