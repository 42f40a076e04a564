"""The token GEMMs of one denoise step (BASELINE configs[1], Bc = 2), libapx against cuBLAS, CUDA-graph replay of 20 back-to-back calls
(so a number includes the launch gap).  APX_GEMM_TRACE=1 prints the in-kernel timeline of CTA 0 instead (test infrastructure).
    PYTHONPATH=. python tools/time_gemm_step.py"""
import os

import torch
import torch.nn.functional as F

from ap_adapter_b200 import _lib

TRACE = os.environ.get("APX_GEMM_TRACE") == "1"


def graph_time(fn, reps=20, iters=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(iters):
            g.replay()
        e1.record(s)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * iters)


CASES = [  # M, N, K, kind, calls per step
    (2000, 2048, 256, "lnfold+geglu", 40), (504, 3072, 384, "lnfold+geglu", 40), (128, 5120, 640, "lnfold+geglu", 48),
    (2000, 256, 1024, "res", 40), (504, 384, 1536, "res", 40), (128, 640, 2560, "res", 48),
    (2000, 768, 256, "lnfold", 40), (2000, 256, 256, "lnfold", 20),
    (504, 1152, 384, "lnfold", 40), (128, 1920, 640, "lnfold", 48),
    (2000, 256, 256, "res", 110), (504, 384, 384, "res", 110), (128, 640, 640, "res", 132),
    (2000, 256, 256, "plain", 0), (504, 384, 384, "plain", 0), (128, 640, 640, "plain", 0),
]


def main():
    torch.manual_seed(0)
    if not TRACE:
        print(f"{'M':>6} {'N':>6} {'K':>6} {'kind':>14} {'n/step':>6} {'apx us':>8} {'torch us':>9}")
    tot_a = tot_t = 0.0
    for m, n, k, kind, calls in CASES:
        a = torch.randn(m, k, device="cuda", dtype=torch.float16)
        w = torch.randn(n, k, device="cuda", dtype=torch.float16) * 0.05
        b16 = torch.randn(n, device="cuda", dtype=torch.float16)
        gamma, beta = torch.rand(k, device="cuda", dtype=torch.float16) + 0.5, torch.randn(k, device="cuda", dtype=torch.float16) * 0.1
        geglu = "geglu" in kind
        res = torch.randn(m, n, device="cuda", dtype=torch.float16) if kind == "res" else None
        if "lnfold" in kind:
            wg, c1, c2 = _lib.fold_layernorm(w, b16, gamma, beta)

            def apx():
                return _lib.linear_lnfold(a, wg, c1, c2, 1e-5, geglu=geglu)

            def ref():
                y = F.linear(F.layer_norm(a, (k,), gamma, beta, 1e-5), w, b16)
                if geglu:   # the reference's GEGLU: chunk, gelu, multiply
                    x, g = y.chunk(2, dim=-1)
                    return x * F.gelu(g)
                return y
        else:
            b32 = b16.float()

            def apx():
                return _lib.linear(a, w, b32, residual=res)

            def ref():
                y = F.linear(a, w, b16)
                return y + res if res is not None else y
        if TRACE:
            apx()
            torch.cuda.synchronize()
            continue
        ta, tt = graph_time(apx), graph_time(ref)
        tot_a += ta * calls
        tot_t += tt * calls
        print(f"{m:6d} {n:6d} {k:6d} {kind:>14} {calls:6d} {ta:8.2f} {tt:9.2f}")
    if not TRACE:
        print(f"weighted by calls per step: libapx {tot_a / 1e3:.2f} ms, torch {tot_t / 1e3:.2f} ms")


if __name__ == "__main__":
    main()
