"""Shared helpers of the GPU parity tests."""
import torch

from oracle import cases, ipattn_oracle, ref_shims

# max|err| / max|ref| against the fp32 oracle fed with the SAME 16-bit-rounded operands.
# north_star: "within 1e-3 relative (fp16/bf16)".  fp16 meets 1e-3 against the fp32 oracle.  bf16 stores 8 significand bits:
# the output rounding alone is 2^-9 = 2e-3 of the value, so NO bf16 implementation can sit within 1e-3 of an fp32 result; the
# bar there is the reference's own bf16 path — the same algorithm run by stock PyTorch ops in bf16 on the same GPU
# (`torch_lowp_site`): ours must be no further from the fp32 oracle than BF16_SLACK x that measured distance (see
# `site_within_bar`), and every measured pair is written to gpurun_out/measured_errors.jsonl -> profiles/.
TOL = {torch.float16: 1e-3, torch.bfloat16: 6e-3}   # bf16 cap: 1.25 x the largest stock-bf16 distance measured (profiles/r2_measured_errors.jsonl: 4.7e-3)
BF16_SLACK = 1.25


def record(name: str, **vals):
    """Append one measured-error record to gpurun_out/measured_errors.jsonl (best effort; the GPU box brings it back)."""
    import json
    import os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "measured_errors.jsonl"), "a") as f:
            f.write(json.dumps(dict(name=name, **{k: (float(v) if isinstance(v, (int, float)) else v) for k, v in vals.items()})) + "\n")
    except OSError:
        pass


def site_within_bar(y, case, dtype, name="", scale=None):
    """The parity bar of one adapter-site output.  fp16: max|err|/max|ref| < 1e-3 against the fp32 oracle.  bf16: no further
    from the fp32 oracle than BF16_SLACK x the stock bf16 PyTorch path's own distance (and in any case < 6e-3)."""
    ref = oracle_site(case, dtype, scale)
    err = rel_err(y, ref)
    if dtype == torch.float16:
        record(name, dtype="f16", err=err, bar=TOL[dtype])
        return err < TOL[dtype], (err,)
    with torch.no_grad():
        err_torch = rel_err(torch_lowp_site(case, dtype, scale), ref)
    record(name, dtype="bf16", err=err, err_torch_bf16=err_torch, bar=max(1e-3, BF16_SLACK * err_torch))
    return err < max(1e-3, BF16_SLACK * err_torch) and err < TOL[dtype], (err, err_torch)


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def rounded(t: torch.Tensor, dtype) -> torch.Tensor:
    return t.to(dtype).float()


def build_site(case: dict, dtype, device="cuda"):
    """(attn stub, product processor) on the device for one oracle case dict."""
    from ap_adapter_b200.attention_processor import IPAttnProcessor2_0

    hidden = case["wq"].shape[0]
    attn = ref_shims.StubAttention(hidden, cases.DCTX, case["heads"])
    proc = IPAttnProcessor2_0(hidden, "test", cases.DCTX, num_tokens=case["num_tokens"], scale=case["scale"])
    with torch.no_grad():
        attn.to_q.weight.copy_(case["wq"]); attn.to_k.weight.copy_(case["wk"]); attn.to_v.weight.copy_(case["wv"])
        attn.to_out[0].weight.copy_(case["wo"]); attn.to_out[0].bias.copy_(case["bo"])
        proc.to_k_ip.weight.copy_(case["wk_ip"]); proc.to_v_ip.weight.copy_(case["wv_ip"])
    return attn.to(device, dtype), proc.to(device, dtype)


def oracle_site(case: dict, dtype, scale=None) -> torch.Tensor:
    r = lambda k: rounded(case[k], dtype)  # noqa: E731
    return ipattn_oracle.ipattn_forward(r("x"), r("ctx"), r("wq"), r("wk"), r("wv"), r("wk_ip"), r("wv_ip"), r("wo"), r("bo"),
                                        case["heads"], case["num_tokens"], case["scale"] if scale is None else scale)


def reference_class_site(case: dict, dtype, scale=None) -> torch.Tensor:
    """The reference's own IPAttnProcessor2_0 (oracle/_ref or /root/reference, unmodified) in fp32 on the CPU, fed the operands
    rounded to `dtype` — what `oracle_site` restates, without the restatement."""
    ap = ref_shims.load_attention_processor()
    hidden = case["wq"].shape[0]
    attn = ref_shims.StubAttention(hidden, cases.DCTX, case["heads"])
    proc = ap.IPAttnProcessor2_0(hidden_size=hidden, name="ref", cross_attention_dim=cases.DCTX, num_tokens=case["num_tokens"],
                                 scale=case["scale"] if scale is None else scale)
    r = lambda k: rounded(case[k], dtype)  # noqa: E731
    with torch.no_grad():
        attn.to_q.weight.copy_(r("wq")); attn.to_k.weight.copy_(r("wk")); attn.to_v.weight.copy_(r("wv"))
        attn.to_out[0].weight.copy_(r("wo")); attn.to_out[0].bias.copy_(r("bo"))
        proc.to_k_ip.weight.copy_(r("wk_ip")); proc.to_v_ip.weight.copy_(r("wv_ip"))
        return proc(attn, r("x"), encoder_hidden_states=r("ctx"))


def torch_lowp_site(case: dict, dtype, scale=None, device="cuda") -> torch.Tensor:
    """The reference algorithm (attention_processor.py:387-459) run by stock PyTorch ops in `dtype` on the GPU — nn.Linear
    + F.scaled_dot_product_attention, i.e. what the reference's own fp16 inference path computes.  Its distance to the
    fp32 oracle is the rounding noise inherent to 16-bit Q/K/V/P storage for this case."""
    import torch.nn.functional as F

    t = lambda k: case[k].to(device, dtype)  # noqa: E731
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
        o = o + (case["scale"] if scale is None else scale) * oa
    return F.linear(o, t("wo"), t("bo"))


def random_site_case(bc, s, c, h, lt, la, scale=0.5, seed=0) -> dict:
    g = torch.Generator().manual_seed(seed)
    n = lambda *shape, std=1.0: torch.randn(*shape, generator=g) * std  # noqa: E731
    d = cases.DCTX
    return dict(x=n(bc, s, c), ctx=n(bc, lt + la, d), wq=n(c, c, std=c ** -0.5), wk=n(c, d, std=1.5 * d ** -0.5),
                wv=n(c, d, std=d ** -0.5), wk_ip=n(c, d, std=1.5 * d ** -0.5), wv_ip=n(c, d, std=d ** -0.5),
                wo=n(c, c, std=c ** -0.5), bo=n(c, std=0.1), heads=h, num_tokens=lt, scale=scale)
