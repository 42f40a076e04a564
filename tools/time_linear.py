"""Graph-replay timing of libapx's tcgen05 GEMM against torch's F.linear at the UNet's token-GEMM shapes (run on the GPU box)."""
import torch
import torch.nn.functional as F
from ap_adapter_b200 import _lib


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


def main():
    torch.manual_seed(0)
    print(f"{'M':>6} {'N':>6} {'K':>6}  {'apx us':>8} {'torch us':>9}  note")
    for m, n, k, note in [(2000, 2048, 256, "ff.proj C=256 (geglu)"), (2000, 256, 1024, "ff.out C=256 (+res)"),
                          (2000, 768, 256, "qkv C=256"), (2000, 256, 256, "to_out C=256"),
                          (500, 3072, 384, "ff.proj C=384"), (500, 384, 1536, "ff.out C=384"), (500, 1152, 384, "qkv C=384"),
                          (500, 384, 384, "to_out C=384"),
                          (126, 5120, 640, "ff.proj C=640"), (126, 640, 2560, "ff.out C=640"), (126, 1920, 640, "qkv C=640"),
                          (126, 640, 640, "to_out C=640"),
                          (8000, 1024, 128, "ff.proj C=128"), (8000, 128, 512, "ff.out C=128"), (8000, 384, 128, "qkv C=128")]:
        a = torch.randn(m, k, device="cuda", dtype=torch.float16)
        w = torch.randn(n, k, device="cuda", dtype=torch.float16) * 0.05
        b16 = torch.randn(n, device="cuda", dtype=torch.float16)
        b32 = b16.float()
        geglu = "geglu" in note
        res = torch.randn(m, n, device="cuda", dtype=torch.float16) if "+res" in note or "to_out" in note else None
        t_apx = graph_time(lambda: _lib.linear(a, w, b32, residual=res, geglu=geglu))
        if geglu:
            def ref():
                x, g = F.linear(a, w, b16).chunk(2, dim=-1)
                return x * F.gelu(g)
        elif res is not None:
            def ref():
                return F.linear(a, w, b16) + res
        else:
            def ref():
                return F.linear(a, w, b16)
        t_ref = graph_time(ref)
        print(f"{m:6d} {n:6d} {k:6d}  {t_apx:8.2f} {t_ref:9.2f}  {note}")


def ff_blocks():
    """The GEGLU FeedForward + residual of a transformer block: two fused libapx GEMMs vs the stock torch graph (5 kernels)."""
    from ap_adapter_b200.unet import FeedForward
    print(f"\n{'C':>5} {'tokens':>7}  {'fused us':>9} {'torch us':>9}")
    for c, tokens in [(256, 1000), (384, 252), (640, 64)]:
        ff = FeedForward(c).eval().half().cuda()
        x = torch.randn(2, tokens, c, device="cuda", dtype=torch.float16)
        with torch.no_grad():
            t_f = graph_time(lambda: ff.fused(x, x))
            t_t = graph_time(lambda: ff(x) + x)
        print(f"{c:5d} {tokens:7d}  {t_f:9.2f} {t_t:9.2f}")


if __name__ == "__main__":
    main()
    ff_blocks()
