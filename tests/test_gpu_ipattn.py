"""GPU parity of the adapter site (IPAttnProcessor2_0 -> apx_ctx_kv + apx_ipattn_fwd through the C ABI) against the
oracle, the committed reference outputs, and size-independent properties at BASELINE's full sizes."""
import os

import numpy as np
import pytest
import torch

from ap_adapter_b200 import _lib
from oracle import cases

from gpu_util import TOL, build_site, oracle_site, random_site_case, record, reference_class_site, rel_err, rounded, site_within_bar, torch_lowp_site
from oracle import ref_shims

pytestmark = pytest.mark.gpu
DTYPES = [torch.float16, torch.bfloat16]


def run_site(case, dtype, scale=None):
    attn, proc = build_site(case, dtype)
    if scale is not None:
        proc.scale = scale
    with torch.no_grad():
        return proc(attn, case["x"].to("cuda", dtype), encoder_hidden_states=case["ctx"].to("cuda", dtype))


def test_extension_is_loaded_and_device_is_blackwell():
    lib = _lib.load()
    assert lib.apx_device_ok() == 1, lib.apx_last_error()
    assert lib.apx_get_impl() == _lib.APX_IMPL_TC


@pytest.mark.parametrize("dtype", DTYPES, ids=["f16", "bf16"])
@pytest.mark.parametrize("name", list(cases.IPATTN_CASES))
def test_site_matches_oracle(name, dtype):
    case = cases.make_ipattn_case(name)
    n0 = _lib.launch_count()
    y = run_site(case, dtype)
    assert _lib.launch_count() > n0, "no libapx kernel was launched"
    ref = oracle_site(case, dtype)
    assert y.shape == ref.shape and y.dtype == dtype
    if name == "site256":
        # The shipped adapter-init weights scaled x8 make the audio softmax sharp (logits of +-30): 16-bit rounding of Q and K
        # alone moves the output by more than 1e-3 — the reference's own PyTorch path in the same dtype (nn.Linear + SDPA on
        # this GPU) sits at the same distance from the fp32 oracle (1.56e-3 fp16 / 1.3e-2 bf16 measured).  The bar for this
        # stress case is therefore "no further from the oracle than the stock 16-bit path", with 10 % slack for summation
        # order.  The same weights exactly as shipped (`site256_real`) go through the plain bar below.
        err = rel_err(y, ref)
        with torch.no_grad():
            err_torch = rel_err(torch_lowp_site(case, dtype), ref)
        record("site_oracle/" + name, dtype=str(dtype), err=err, err_torch=err_torch)
        assert err < max(1e-3, 1.1 * err_torch), (err, err_torch)
    else:
        ok, errs = site_within_bar(y, case, dtype, "site_oracle/" + name)
        assert ok, errs


@pytest.mark.skipif(not ref_shims.reference_processors_available(), reason="oracle/_ref not built (oracle/build_ref.py)")
@pytest.mark.parametrize("name", ["cfg1", "site256_real", "site384", "site640"])
def test_site_matches_the_reference_class_itself(name):
    """libapx against the reference's OWN IPAttnProcessor2_0 class (oracle/_ref: its attention_processor module, compiled from the unmodified source), run in
    fp32 on the CPU of this box on the same fp16-rounded operands — no restatement in between.  Bar: the north star's 1e-3
    (the real-weight site is held to the stock fp16 path's own distance, as in test_site_matches_oracle)."""
    if name not in cases.IPATTN_CASES:
        pytest.skip(f"no case {name}")
    case = cases.make_ipattn_case(name)
    y = run_site(case, torch.float16)
    ref = reference_class_site(case, torch.float16)
    err = rel_err(y, ref)
    err_port = rel_err(oracle_site(case, torch.float16), ref)   # the restatement against the class it restates, same inputs
    record("site_reference_class/" + name, dtype="f16", err=err, port_vs_class=err_port)
    assert err_port < 1e-5, err_port
    if name == "site256_real":
        with torch.no_grad():
            err_torch = rel_err(torch_lowp_site(case, torch.float16), ref)
        assert err < max(1e-3, 1.1 * err_torch), (err, err_torch)
    else:
        assert err < TOL[torch.float16], err


@pytest.mark.parametrize("name", list(cases.IPATTN_CASES))
def test_site_matches_reference_output_fixture(name):
    """Against outputs of the reference's own IPAttnProcessor2_0 (fp32, unrounded operands): the fp16 operand rounding
    is now part of the error, hence 3e-3."""
    gold = np.load(os.path.join(cases.GOLDEN_DIR, "ipattn_reference.npz"))[name]
    y = run_site(cases.make_ipattn_case(name), torch.float16)
    y = y.reshape(-1, y.shape[-1])[:: cases.ROW_STRIDE.get(name, 1)]
    assert rel_err(y, torch.from_numpy(gold)) < 3e-3


@pytest.mark.parametrize("shape", [
    (1, 1, 256, 8, 8, 8),        # single query row
    (3, 129, 256, 8, 8, 24),     # one row into the second 128-row tile, La not a multiple of 16
    (2, 64, 640, 8, 8, 128),     # d = 80
    (2, 252, 384, 8, 8, 32),     # d = 48
    (1, 70, 320, 8, 77, 8),      # d = 40 (config 1 head size), long text segment
    (2, 100, 256, 8, 0, 40),     # no text tokens: audio-only context
    (2, 100, 256, 8, 5, 0),      # no audio tokens
    (1, 200, 256, 8, 8, 512),    # un-pooled AudioMAE grid
    (2, 96, 512, 16, 8, 32),     # 16 heads, d = 32
    (16, 252, 384, 8, 8, 128),   # BASELINE batch 8 (Bc = 16) at the C = 384 sites
    (16, 64, 640, 8, 8, 128),    # ... and at the C = 640 sites
    (4, 64, 640, 8, 8, 32),      # style_transfer preset (pool 4x4)
    (8, 1000, 256, 8, 8, 8),     # pool 8x8
], ids=lambda s: "x".join(map(str, s)))
def test_site_edge_shapes(shape):
    case = random_site_case(*shape, scale=0.7, seed=sum(shape))
    y = run_site(case, torch.float16)
    assert rel_err(y, oracle_site(case, torch.float16)) < TOL[torch.float16]


@pytest.mark.parametrize("dtype", DTYPES, ids=["f16", "bf16"])
@pytest.mark.parametrize("name", ["cfg1", "site256_real", "site384", "site640", "ctx2d", "text_only"])
def test_one_head_site_kernel_matches_oracle(name, dtype):
    """The one-head-per-CTA (cluster) site kernel pinned on every shape it covers — C = 320 (d = 40, padded to 48) / 256 / 384 / 640,
    with and without audio tokens — at the same bar as the product's own choice of kernel."""
    lib = _lib.load()
    case = cases.make_ipattn_case(name)
    bc, s, c = (case["x"].shape if case["x"].dim() == 3 else (1,) + tuple(case["x"].shape))
    lctx = case["ctx"].shape[-2]
    lt = min(case["num_tokens"], lctx)
    assert lib.apx_set_site_variant(_lib.APX_SITE_ONEHEAD) == 0
    try:
        assert lib.apx_ipattn_site_kernel(bc, s, c, case["heads"], lt, lctx - lt) == 2
        n0 = _lib.launch_count()
        y = run_site(case, dtype)
        # ctx K/V build (4 GEMMs + merge + pack) + weight packs + ONE site launch: the site itself must not be three kernels
        assert _lib.launch_count() - n0 <= 4 + 2 + 2 + 1
        ok, errs = site_within_bar(y, case, dtype, "site_onehead/" + name)
    finally:
        assert lib.apx_set_site_variant(_lib.APX_SITE_AUTO) == 0
    assert ok, errs


@pytest.mark.parametrize("dtype", DTYPES, ids=["f16", "bf16"])
def test_four_warpgroup_site_kernel_matches_oracle(dtype):
    """The opt-in all-heads variant with four softmax warpgroups (APX_SITE_ALLHEADS4; score rows split over two threads, audio maxima
    exchanged through shared memory) on the S = 1000, C = 256 site with the shipped adapter weights."""
    lib = _lib.load()
    case = cases.make_ipattn_case("site256_real")
    assert lib.apx_set_site_variant(_lib.APX_SITE_ALLHEADS4) == 0
    try:
        assert lib.apx_ipattn_site_kernel(2, 1000, 256, 8, 8, 128) == 3
        y = run_site(case, dtype)
        ok, errs = site_within_bar(y, case, dtype, "site_4wg/site256_real")
    finally:
        assert lib.apx_set_site_variant(_lib.APX_SITE_AUTO) == 0
    assert ok, errs


@pytest.mark.parametrize("dtype", DTYPES, ids=["f16", "bf16"])
@pytest.mark.parametrize("shape", [(2, 1000, 256, 8, 8, 128), (16, 1000, 256, 8, 8, 128), (3, 129, 256, 8, 8, 24), (2, 200, 128, 4, 8, 32)],
                         ids=lambda s: "x".join(map(str, s)))
def test_site_with_folded_norm2_matches_oracle(shape, dtype):
    """apx_ipattn_fwd_ln: the block's norm2 LayerNorm inside the all-heads site kernel (row statistics from the staged X tiles, Wq *
    gamma in the pack, correction in the Q conversion) against LayerNorm -> oracle site in fp32 on the same 16-bit operands, with and
    without the residual folded in (Y aliasing X)."""
    bc, s, c, h, lt, la = shape
    case = random_site_case(bc, s, c, h, lt, la, scale=0.6, seed=sum(shape))
    g = torch.Generator().manual_seed(1)
    case["x"] = case["x"] * 1.7 + 0.6 * torch.randn(bc, s, 1, generator=g) + 0.25           # rows with their own offset and scale
    norm = torch.nn.LayerNorm(c)
    with torch.no_grad():
        norm.weight.normal_(1.0, 0.25, generator=g); norm.bias.normal_(0.0, 0.2, generator=g)
    attn, proc = build_site(case, dtype)
    import copy
    norm_d = copy.deepcopy(norm).to("cuda", dtype)
    x = case["x"].to("cuda", dtype)
    ctx = case["ctx"].to("cuda", dtype)
    assert not proc.can_fold_norm(attn, x, norm_d, ctx)          # autograd on: training keeps the separate LayerNorm
    with torch.no_grad():
        assert proc.can_fold_norm(attn, x, norm_d, ctx)
        n0 = _lib.launch_count()
        y = proc(attn, x, encoder_hidden_states=ctx, fold_norm=norm_d)
        built = _lib.launch_count() - n0
        n0 = _lib.launch_count()
        h_res = x.clone()
        y2 = proc(attn, x, encoder_hidden_states=ctx, fold_norm=norm_d, residual_inout=h_res)      # h + site(norm2(h)), in place
        assert _lib.launch_count() - n0 == 1 and built >= 1 and y2.data_ptr() == h_res.data_ptr()
        h_alias = x.clone()
        y3 = proc(attn, h_alias, encoder_hidden_states=ctx, fold_norm=norm_d, residual_inout=h_alias)  # X and Y are the same buffer
    xr = case["x"].to(dtype).float()
    xn = torch.nn.functional.layer_norm(xr, (c,), norm.weight.to(dtype).float(), norm.bias.to(dtype).float(), norm.eps)
    ref = oracle_site(dict(case, x=xn), dtype)          # oracle_site rounds x again: the reference's LayerNorm output is 16-bit too
    err = rel_err(y, ref)
    record("site_ln/" + "x".join(map(str, shape)), dtype=str(dtype), err=err)
    assert err < (1.5e-3 if dtype == torch.float16 else 1e-2)
    assert torch.equal(y2, (y.float() + x.float()).to(dtype)) or rel_err(y2, ref + xr) < (1.5e-3 if dtype == torch.float16 else 1e-2)
    assert torch.equal(y3, y2)


def test_tensor_core_and_cuda_core_paths_agree():
    lib = _lib.load()
    case = random_site_case(2, 300, 384, 8, 8, 128, seed=5)
    try:
        lib.apx_set_impl(_lib.APX_IMPL_SIMT)
        y_simt = run_site(case, torch.float16)
    finally:
        lib.apx_set_impl(_lib.APX_IMPL_TC)
    y_tc = run_site(case, torch.float16)
    assert rel_err(y_tc, y_simt) < 1e-3
    assert rel_err(y_simt, oracle_site(case, torch.float16)) < 1e-3


def test_full_size_properties_batch8():
    """BASELINE batch 8 (Bc = 16) at the S = 1000, C = 256 site, La = 128: too big for the CPU oracle in seconds, so check
    properties the algorithm guarantees: determinism, batch independence (bitwise), linearity in ap_scale, invariance
    to a permutation of the audio tokens, and oracle agreement on a slice."""
    case = random_site_case(16, 1000, 256, 8, 8, 128, seed=11)
    y0 = run_site(case, torch.float16, scale=0.0)
    y5 = run_site(case, torch.float16, scale=0.5)
    y5b = run_site(case, torch.float16, scale=0.5)
    y1 = run_site(case, torch.float16, scale=1.0)
    assert torch.equal(y5, y5b)
    lin = (y1.float() - y0.float()) - 2 * (y5.float() - y0.float())
    assert lin.abs().max().item() < 4e-3 * y1.float().abs().max().item()
    # batch element 3 alone == batch element 3 inside the batch, bit for bit (the context GEMMs pick the same split-K factor at
    # both sizes and sum the partials in rank order, so a row's value does not depend on its position)
    sub = dict(case, x=case["x"][3:4], ctx=case["ctx"][3:4])
    assert torch.equal(run_site(sub, torch.float16, scale=0.5)[0], y5[3])
    # permuting the audio tokens only reorders a sum
    perm = torch.cat([torch.arange(8), 8 + torch.randperm(128, generator=torch.Generator().manual_seed(0))])
    yp = run_site(dict(case, ctx=case["ctx"][:, perm]), torch.float16, scale=0.5)
    assert rel_err(yp, y5) < 1e-3
    # oracle on two batch elements
    sl = dict(case, x=case["x"][[0, 15]], ctx=case["ctx"][[0, 15]])
    assert rel_err(y5[[0, 15]], oracle_site(sl, torch.float16, scale=0.5)) < TOL[torch.float16]


def test_context_cache_follows_weights_and_context():
    case = random_site_case(2, 64, 256, 8, 8, 16, seed=3)
    attn, proc = build_site(case, torch.float16)
    x, ctx = case["x"].cuda().half(), case["ctx"].cuda().half()
    with torch.no_grad():
        y_a = proc(attn, x, encoder_hidden_states=ctx)
        blob = proc._ctx_cache[1]
        y_b = proc(attn, x, encoder_hidden_states=ctx)
        assert proc._ctx_cache[1] is blob and torch.equal(y_a, y_b)  # reused
        proc.to_v_ip.weight.mul_(2.0)  # in-place update bumps ._version -> rebuild
        y_c = proc(attn, x, encoder_hidden_states=ctx)
        assert proc._ctx_cache[1] is not blob
        case2 = dict(case, wv_ip=case["wv_ip"].half().float() * 2.0)
        assert rel_err(y_c, oracle_site(case2, torch.float16)) < 2e-3
        ctx2 = ctx.clone()
        ctx2[:, 8:] = 0
        y_d = proc(attn, x, encoder_hidden_states=ctx2)
        assert rel_err(y_d, oracle_site(dict(case2, ctx=ctx2.float().cpu()), torch.float16)) < 2e-3
        # reference-style weight replacement (inference.py:56-57)
        proc.to_k_ip.weight = torch.nn.Parameter(torch.zeros_like(proc.to_k_ip.weight))
        y_e = proc(attn, x, encoder_hidden_states=ctx)
        assert rel_err(y_e, oracle_site(dict(case2, wk_ip=torch.zeros_like(case["wk_ip"])), torch.float16)) < 2e-3


def test_two_dimensional_context_is_unsqueezed():
    case = cases.make_ipattn_case("ctx2d")
    assert case["ctx"].dim() == 2
    assert rel_err(run_site(case, torch.float16), oracle_site(case, torch.float16)) < 1e-3


def test_unsupported_inputs_raise_instead_of_falling_back():
    case = random_site_case(1, 32, 256, 8, 8, 8)
    attn, proc = build_site(case, torch.float16)
    x, ctx = case["x"].cuda().half(), case["ctx"].cuda().half()
    with torch.no_grad():
        with pytest.raises(_lib.ApxError, match="dtype"):
            proc(attn.float(), x.float(), encoder_hidden_states=ctx.float())
        attn.half()
        with pytest.raises(NotImplementedError):
            proc(attn, x, encoder_hidden_states=ctx, attention_mask=torch.zeros(1, 1, 16, device="cuda"))
        with pytest.raises(_lib.ApxError, match="batch"):
            proc(attn, x, encoder_hidden_states=ctx.repeat(2, 1, 1))
    # training needs the audio branch: a text-only context under autograd is an error, not a silent zero gradient
    with pytest.raises(_lib.ApxError, match="audio tokens"):
        proc(attn, x.requires_grad_(), encoder_hidden_states=ctx[:, :8])
    lib = _lib.load()
    # C ABI level: head dim 20 is not a multiple of 8
    buf = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    rc = lib.apx_ipattn_fwd(_lib.ptr(buf), _lib.ptr(buf), _lib.ptr(buf), _lib.ptr(buf), None, None, _lib.ptr(buf), 1, 16, 160, 8, 8, 8,
                            0.5, 0, _lib.ptr(buf), buf.numel(), None)
    assert rc == 2 and b"head dim" in lib.apx_last_error()
    rc = lib.apx_ipattn_fwd(_lib.ptr(buf), _lib.ptr(buf), _lib.ptr(buf), _lib.ptr(buf), None, None, _lib.ptr(buf), 1, 16, 256, 8, 8, 8,
                            0.5, 0, _lib.ptr(buf), 16, None)
    assert rc == 4  # workspace too small


@pytest.mark.parametrize("dtype", DTYPES, ids=["f16", "bf16"])
def test_linear_epilogues(dtype):
    g = torch.Generator().manual_seed(0)
    m, n, k = 333, 384, 768
    a = torch.randn(m, k, generator=g).to(dtype).cuda()
    w = (torch.randn(n, k, generator=g) * k ** -0.5).to(dtype).cuda()
    b = torch.randn(n, generator=g).cuda()
    r = torch.randn(m, n, generator=g).cuda()
    ref = a.float() @ w.float().t() + b
    assert rel_err(_lib.linear(a, w, b, out_dtype=torch.float32), ref) < 1e-5 * 50
    assert rel_err(_lib.linear(a, w, b, gelu=True), torch.nn.functional.gelu(ref)) < TOL[dtype]
    assert rel_err(_lib.linear(a, w, b, residual=r, out_dtype=torch.float32), ref + r) < 1e-5 * 50
    with pytest.raises(_lib.ApxError, match="multiple of 64"):
        _lib.linear(a[:, :100].contiguous(), w[:, :100].contiguous())


def test_segmented_attention_kernel():
    g = torch.Generator().manual_seed(1)
    b, s, h, d, l0, l = 2, 77, 12, 64, 13, 50
    q, k, v = (torch.randn(b, n, h * d, generator=g).half().cuda() for n in (s, l, l))
    o = _lib.attention(q, k, v, h, l0=l0, w0=1.0, w1=0.3)

    def sdpa(kk, vv):
        qh, kh, vh = (t.float().view(b, -1, h, d).transpose(1, 2) for t in (q, kk, vv))
        return torch.nn.functional.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(b, s, h * d)

    ref = sdpa(k[:, :l0], v[:, :l0]) + 0.3 * sdpa(k[:, l0:], v[:, l0:])
    assert rel_err(o, ref) < 1e-3


# ---- adapter training (BASELINE config 5): backward through the site --------------------------------------
@pytest.mark.parametrize("dtype", DTYPES, ids=["f16", "bf16"])
@pytest.mark.parametrize("shape", [
    (2, 200, 256, 8, 8, 32),     # site kernel geometry (cluster launch), La = 32
    (1, 130, 384, 8, 8, 128),    # three-kernel forward, d = 48, ragged query tile
    (2, 64, 640, 8, 8, 8),       # d = 80
    (1, 70, 256, 8, 0, 24),      # no text tokens
    (2, 150, 256, 8, 8, 512),    # un-pooled AudioMAE grid (pooling rate 1, drawn 25 % of the time: train_apadapter_v2.py:445-446)
    (1, 100, 384, 8, 8, 512),
    (2, 64, 640, 8, 8, 512),
], ids=["256_la32", "384_la128", "640_la8", "audio_only", "256_la512", "384_la512", "640_la512"])
def test_adapter_backward_matches_autograd_of_the_oracle(shape, dtype):
    """dX, d to_k_ip.weight, d to_v_ip.weight from apx_ipattn_bwd against torch autograd through the fp32 oracle evaluated on
    the same 16-bit-rounded operands, for a random upstream gradient.  The backward chains five 16-bit roundings (Q, dO, P
    recomputed from 16-bit K, dQ, dX); measured max over the seven shapes: 9.4e-4 (fp16) / 6.1e-3 (bf16)
    (profiles/r2_measured_errors.jsonl) -> the bar is 2x that: 2e-3 / 1.2e-2."""
    from oracle import ipattn_oracle

    bc, s, c, h, lt, la = shape
    case = random_site_case(bc, s, c, h, lt, la, scale=1.0, seed=3)   # training runs with scale = 1.0 (train_apadapter_v2.py:644)
    attn, proc = build_site(case, dtype)
    for p in attn.parameters():
        p.requires_grad_(False)
    x = case["x"].to("cuda", dtype).requires_grad_(True)
    enc = case["ctx"].to("cuda", dtype)
    g = torch.Generator().manual_seed(11)
    dy = torch.randn(bc, s, c, generator=g)
    n0 = _lib.launch_count()
    y = proc(attn, x, encoder_hidden_states=enc)
    y.backward(dy.to("cuda", dtype))
    assert _lib.launch_count() > n0
    assert x.grad is not None and proc.to_k_ip.weight.grad is not None and proc.to_v_ip.weight.grad is not None

    r = lambda k: case[k].to(dtype).float()  # noqa: E731
    xr = r("x").requires_grad_(True)
    wk_ip, wv_ip = r("wk_ip").requires_grad_(True), r("wv_ip").requires_grad_(True)
    ref = ipattn_oracle.ipattn_forward(xr, r("ctx"), r("wq"), r("wk"), r("wv"), wk_ip, wv_ip, r("wo"), r("bo"), h, lt, 1.0)
    ref.backward(dy.to(dtype).float())
    assert rel_err(y, ref) < TOL[dtype]
    tol = 2e-3 if dtype == torch.float16 else 1.2e-2
    record("site_bwd/" + "x".join(map(str, shape)), dtype=str(dtype), dx=rel_err(x.grad, xr.grad),
           dwk=rel_err(proc.to_k_ip.weight.grad, wk_ip.grad), dwv=rel_err(proc.to_v_ip.weight.grad, wv_ip.grad), bar=tol)
    assert rel_err(x.grad, xr.grad) < tol
    assert rel_err(proc.to_k_ip.weight.grad, wk_ip.grad) < tol
    assert rel_err(proc.to_v_ip.weight.grad, wv_ip.grad) < tol


def test_adapter_training_step_reduces_the_loss():
    """Three AdamW steps on the two adapter weights of one site (everything else frozen) lower an MSE loss: the gradients point
    the right way and the context cache follows the weight updates."""
    case = random_site_case(2, 128, 256, 8, 8, 32, scale=1.0, seed=5)
    attn, proc = build_site(case, torch.bfloat16)
    proc = proc.float()                       # fp32 master weights, bf16 compute (train_apadapter_v2.py:645)
    for p in attn.parameters():
        p.requires_grad_(False)
    opt = torch.optim.AdamW(proc.parameters(), lr=1e-2)
    x = case["x"].to("cuda", torch.bfloat16)
    enc = case["ctx"].to("cuda", torch.bfloat16)
    target = torch.randn(2, 128, 256, generator=torch.Generator().manual_seed(1)).to("cuda")
    losses = []
    for _ in range(4):
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(proc(attn, x, encoder_hidden_states=enc).float(), target)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("c,heads,s", [(256, 8, 1000), (256, 8, 77), (384, 8, 252), (640, 8, 64)])
def test_residual_folded_into_the_site_matches_separate_add(c, heads, s, dtype):
    """apx_ipattn_fwd_add (§8f-2, "residual add -> epilogue"): h += site(norm2(h)) in place equals the caller's elementwise add of the
    same site output, bit for bit (one 16-bit rounding of the site, one of the sum) — fused kernel (C=256) and three-kernel form."""
    from ap_adapter_b200.attention_processor import IPAttnProcessor2_0
    from ap_adapter_b200.unet import Attention
    torch.manual_seed(c + s)
    attn = Attention(c, 768, heads, c // heads).to("cuda", dtype).eval()
    proc = IPAttnProcessor2_0(hidden_size=c, cross_attention_dim=768, num_tokens=8, scale=0.5, name="t").to("cuda", dtype)
    attn.set_processor(proc)
    x = torch.randn(2, s, c, device="cuda", dtype=dtype)
    h = torch.randn(2, s, c, device="cuda", dtype=dtype)
    ctx = torch.randn(2, 8 + 128, 768, device="cuda", dtype=dtype)
    with torch.no_grad():
        y = attn(x, encoder_hidden_states=ctx)
        want = y + h
        h2 = h.clone()
        got = attn(x, encoder_hidden_states=ctx, residual_inout=h2)
    assert got.data_ptr() == h2.data_ptr()          # in place
    assert torch.equal(got, want)
