#!/usr/bin/env python
"""Error table of the adapter site on the GPU (run under gpurun): for every oracle case and dtype, max|err|/max|ref| of
  ours (fused product path), ours (three tcgen05 kernels), and the reference algorithm run by PyTorch in the same 16-bit
  dtype on the GPU (nn.Linear + F.scaled_dot_product_attention: the stock reference path),
all against the fp32 CPU oracle fed the same 16-bit-rounded operands.  TEST INFRASTRUCTURE (imports oracle/)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from ap_adapter_b200 import _lib  # noqa: E402
from gpu_util import build_site, oracle_site, random_site_case, rel_err  # noqa: E402
from oracle import cases  # noqa: E402


def torch_lowp_path(case, dtype):
    """IPAttnProcessor2_0.__call__ (attention_processor.py:387-459) with torch ops in `dtype` on the GPU."""
    dev = "cuda"
    t = lambda k: case[k].to(dev, dtype)  # noqa: E731
    x, ctx = t("x"), t("ctx")
    if ctx.dim() < 3:
        ctx = ctx.unsqueeze(0)
    h, nt = case["heads"], case["num_tokens"]
    b = ctx.shape[0]
    q = F.linear(x, t("wq"))
    c = q.shape[-1]
    sp = lambda z: z.view(b, -1, h, c // h).transpose(1, 2)  # noqa: E731
    text, audio = ctx[:, :nt], ctx[:, nt:]
    o = torch.zeros_like(q)
    if text.shape[1]:
        o = F.scaled_dot_product_attention(sp(q), sp(F.linear(text, t("wk"))), sp(F.linear(text, t("wv")))).transpose(1, 2).reshape(b, -1, c)
    if audio.shape[1]:
        oa = F.scaled_dot_product_attention(sp(q), sp(F.linear(audio, t("wk_ip"))), sp(F.linear(audio, t("wv_ip")))).transpose(1, 2).reshape(b, -1, c)
        o = o + case["scale"] * oa
    return F.linear(o, t("wo"), t("bo"))


def main():
    lib = _lib.load()
    todo = [(n, cases.make_ipattn_case(n)) for n in cases.IPATTN_CASES]
    for shape in [(3, 129, 256, 8, 8, 24), (2, 252, 384, 8, 8, 32), (2, 252, 384, 8, 8, 128), (2, 1000, 256, 8, 8, 128), (2, 64, 640, 8, 8, 128),
                  (1, 70, 320, 8, 77, 8), (2, 96, 512, 16, 8, 32)]:
        todo.append(("rand_" + "x".join(map(str, shape)), random_site_case(*shape, scale=0.7, seed=sum(shape))))
    print(f"{'case':34s} {'dtype':5s} {'fused?':6s} {'ours':>10s} {'unfused':>10s} {'torch-lowp':>10s}")
    for name, case in todo:
        for dtype in (torch.float16, torch.bfloat16):
            ref = oracle_site(case, dtype)
            c, h = case["wq"].shape[0], case["heads"]
            lt = case["num_tokens"]
            la = case["ctx"].shape[-2] - lt
            fused = lib.apx_ipattn_fused_ok(c, h, lt, la)
            out = {}
            for tag, impl in (("ours", _lib.APX_IMPL_TC), ("unfused", _lib.APX_IMPL_TC_UNFUSED)):
                lib.apx_set_impl(impl)
                attn, proc = build_site(case, dtype)
                with torch.no_grad():
                    out[tag] = rel_err(proc(attn, case["x"].to("cuda", dtype), encoder_hidden_states=case["ctx"].to("cuda", dtype)), ref)
            lib.apx_set_impl(_lib.APX_IMPL_TC)
            with torch.no_grad():
                out["torch"] = rel_err(torch_lowp_path(case, dtype), ref)
            print(f"{name:34s} {str(dtype)[6:]:5s} {fused:6d} {out['ours']:10.3e} {out['unfused']:10.3e} {out['torch']:10.3e}", flush=True)


if __name__ == "__main__":
    main()
