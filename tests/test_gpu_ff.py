"""§8(f-2) neighbour of the hot path: the transformer block's GEGLU FeedForward + residual as two tcgen05 GEMMs with fused
epilogues (APX_EPI_GEGLU, APX_EPI_RES16), against the unfused fp32 torch graph of the same module
(diffusers FeedForward(activation_fn="geglu"), which AudioLDM2's BasicTransformerBlock ends with)."""
import pytest
import torch
import torch.nn.functional as F

from ap_adapter_b200 import _lib
from ap_adapter_b200.unet import BasicTransformerBlock, FeedForward

from gpu_util import TOL, rel_err, rounded

pytestmark = pytest.mark.gpu


def _ff_ref(ff: FeedForward, x: torch.Tensor, res: torch.Tensor, dtype) -> torch.Tensor:
    """fp32 math on the 16-bit-rounded operands; the hidden activation is rounded to `dtype` like the kernel's output."""
    w1, b1 = ff.net[0].proj.weight.float(), ff.net[0].proj.bias.float()
    w2, b2 = ff.net[2].weight.float(), ff.net[2].bias.float()
    a, gate = F.linear(x.float(), w1, b1).chunk(2, dim=-1)
    hidden = rounded(a * F.gelu(gate), dtype)
    return F.linear(hidden, w2, b2) + res.float()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("dim,tokens", [(128, 4000), (256, 1000), (384, 250), (640, 77), (256, 1), (640, 8)])
def test_fused_feedforward_matches_fp32(dim, tokens, dtype):
    torch.manual_seed(dim + tokens)
    ff = FeedForward(dim).eval().to("cuda", dtype)
    x = torch.randn(2, tokens, dim, device="cuda", dtype=dtype)
    res = torch.randn(2, tokens, dim, device="cuda", dtype=dtype)
    n0 = _lib.launch_count()
    with torch.no_grad():
        assert ff.fused_ok(x)
        y = ff.fused(x, res)
    assert _lib.launch_count() - n0 == 2
    assert y.shape == x.shape and y.dtype == dtype
    assert rel_err(y, _ff_ref(ff, x, res, dtype)) < TOL[dtype]


def test_geglu_epilogue_agrees_with_cuda_core_gemm():
    """The same epilogue flags through the cross-check GEMM (APX_IMPL_SIMT)."""
    torch.manual_seed(3)
    a = torch.randn(300, 128, device="cuda", dtype=torch.float16)
    w = torch.randn(512, 128, device="cuda", dtype=torch.float16) * 0.1
    b = torch.randn(512, device="cuda")
    r = torch.randn(300, 256, device="cuda", dtype=torch.float16)
    y_tc = _lib.linear(a, w, b, residual=r, geglu=True)
    lib = _lib.load()
    prev = lib.apx_get_impl()
    lib.apx_set_impl(1)
    try:
        y_simt = _lib.linear(a, w, b, residual=r, geglu=True)
    finally:
        lib.apx_set_impl(prev)
    z = F.linear(a.float(), w.float(), b)
    ref = z[:, 0::2] * F.gelu(z[:, 1::2]) + r.float()
    assert rel_err(y_tc, ref) < 1e-3 and rel_err(y_simt, ref) < 1e-3


def test_block_takes_fused_route_and_matches_unfused():
    torch.manual_seed(5)
    blk = BasicTransformerBlock(256, 4, 64, cross_attention_dim=None, double_self_attention=False).eval().half().cuda()
    h = torch.randn(2, 400, 256, device="cuda", dtype=torch.float16)      # 800 rows: below the Q(KV) LayerNorm-fold threshold
    with torch.no_grad():
        BasicTransformerBlock.fold_residual = True
        add_ln = BasicTransformerBlock.add_ln
        try:
            n0 = _lib.launch_count()
            y = blk(h)
            # norm1 as libapx's LayerNorm (apx_add_layernorm without an addend), attn1's self-attention core (400 keys, d = 64) and its
            # to_out + residual GEMM (opt-in), the two FeedForward GEMMs
            assert _lib.launch_count() - n0 == 4 + (1 if BasicTransformerBlock.add_ln else 0)
            FeedForward.use_fused = BasicTransformerBlock.fold_residual = False
            BasicTransformerBlock.add_ln = False
            y_ref = blk(h)
        finally:
            FeedForward.use_fused, BasicTransformerBlock.fold_residual = True, "ip"
            BasicTransformerBlock.add_ln = add_ln
    assert rel_err(y, y_ref) < 2e-3
    # grad mode keeps the differentiable torch graph
    n0 = _lib.launch_count()
    blk(h)
    assert _lib.launch_count() == n0


def test_repack_after_weight_update():
    ff = FeedForward(128).eval().half().cuda()
    x = torch.randn(1, 64, 128, device="cuda", dtype=torch.float16)
    with torch.no_grad():
        y0 = ff.fused(x, x)
        ff.net[0].proj.weight.mul_(0.5)
        ff.net[0].proj.bias.mul_(0.5)
        y1 = ff.fused(x, x)
        FeedForward.use_fused = False
        try:
            ref = ff(x) + x
        finally:
            FeedForward.use_fused = True
    assert not torch.equal(y0, y1)
    assert rel_err(y1, ref) < 2e-3


@pytest.mark.parametrize("c,heads,s,cross", [(256, 8, 1000, None), (384, 8, 252, 1024), (640, 8, 64, None)])
def test_plain_site_folds_to_out_and_residual_into_one_gemm(c, heads, s, cross):
    """AttnProcessor2_0 with `residual_inout`: to_out + bias + the block's residual add as one libapx GEMM (APX_EPI_RES16)
    against the stock F.linear + elementwise add of the same processor (reference attention_processor.py:280-292 + the
    caller's add)."""
    from ap_adapter_b200.attention_processor import AttnProcessor2_0
    from ap_adapter_b200.unet import Attention
    torch.manual_seed(c)
    attn = Attention(c, cross, heads, c // heads).half().cuda().eval()
    attn.set_processor(AttnProcessor2_0())
    x = torch.randn(2, s, c, device="cuda", dtype=torch.float16)
    h = torch.randn(2, s, c, device="cuda", dtype=torch.float16)
    ctx = None if cross is None else torch.randn(2, 16, cross, device="cuda", dtype=torch.float16)
    with torch.no_grad():
        want = attn(x, encoder_hidden_states=ctx) + h
        n0 = _lib.launch_count()
        h_before = h.clone()
        got = attn(x, encoder_hidden_states=ctx, residual_inout=h)
        proc = attn.processor
        flash = (ctx is None and s >= proc.flash_min_keys and c // heads in proc.flash_head_dims
                 and 2 * heads * ((s + 127) // 128) <= proc.flash_max_ctas)
        assert _lib.launch_count() - n0 == 1 + int(flash)       # the to_out + residual GEMM (+ the libapx self-attention core)
    assert torch.equal(h, h_before) and got.data_ptr() != h.data_ptr()   # the plain sites do not write into the residual stream
    assert rel_err(got, want) < 1e-3
    # grad mode: same signature, stock kernels
    n0 = _lib.launch_count()
    got2 = attn(x, encoder_hidden_states=ctx, residual_inout=h)
    assert _lib.launch_count() == n0 and rel_err(got2, want) < 1e-3


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
@pytest.mark.parametrize("m,k,n,geglu,res", [(2000, 256, 2048, True, False), (504, 384, 3072, True, False), (128, 640, 5120, True, False),
                                             (333, 256, 768, False, True), (1026, 768, 2304, False, False),
                                             (128, 640, 1920, False, False), (100, 640, 640, False, True), (70, 1024, 256, True, True),
                                             (2500, 256, 1024, False, True)],
                         ids=["ff256", "ff384", "ff640", "qkv_res", "vit_qkv", "qkv640_splitk", "q640_splitk_res", "geglu_splitk_res",
                              "wide_res"])
def test_linear_with_folded_layernorm_matches_torch(m, k, n, geglu, res, dtype):
    """apx_linear_lnfold (LayerNorm of the A rows folded into the tcgen05 GEMM: statistics from the A tiles in shared memory, correction
    in the epilogue) against LayerNorm -> Linear (-> GEGLU) in fp32 on the same 16-bit operands."""
    import torch.nn.functional as F
    from ap_adapter_b200 import _lib
    from gpu_util import record, rel_err
    g = torch.Generator().manual_seed(m + n)
    x = (torch.randn(m, k, generator=g) * 1.5 + 0.8 * torch.randn(m, 1, generator=g) + 0.3).to(dtype)      # rows with their own offsets
    w = (torch.randn(n, k, generator=g) * k ** -0.5).to(dtype)
    b = torch.randn(n, generator=g) * 0.2
    gamma, beta = 1.0 + 0.3 * torch.randn(k, generator=g), 0.2 * torch.randn(k, generator=g)
    r = torch.randn(m, n // 2 if geglu else n, generator=g).to(dtype) if res else None
    ref = F.linear(F.layer_norm(x.float(), (k,), gamma, beta, 1e-5), w.float(), b)
    if geglu:
        ref = ref[:, 0::2] * F.gelu(ref[:, 1::2])
    if res:
        ref = ref.to(dtype).float() + r.float()
    wg, c1, c2 = _lib.fold_layernorm(w.cuda(), b.cuda(), gamma.cuda(), beta.cuda())
    n0 = _lib.launch_count()
    y = _lib.linear_lnfold(x.cuda(), wg, c1, c2, 1e-5, residual=r.cuda() if res else None, geglu=geglu)
    assert _lib.launch_count() - n0 == 1
    err = rel_err(y, ref)
    record(f"linear_lnfold/{m}x{k}x{n}", dtype=str(dtype), err=err)
    assert err < (2e-3 if dtype == torch.float16 else 1.2e-2)      # 16-bit rounding of W * gamma and of the output


def test_feedforward_with_folded_norm_matches_separate_layernorm():
    """FeedForward.fused_ln(h, norm3) == FeedForward.fused(norm3(h), h) up to the 16-bit rounding points."""
    from ap_adapter_b200.unet import FeedForward
    from gpu_util import rel_err
    torch.manual_seed(3)
    ff = FeedForward(384).cuda().half().eval()
    norm = torch.nn.LayerNorm(384).cuda().half()
    with torch.no_grad():
        norm.weight.normal_(1.0, 0.2); norm.bias.normal_(0.0, 0.2)
        h = torch.randn(2, 252, 384, device="cuda", dtype=torch.float16)
        a = ff.fused_ln(h, norm)
        b = ff.fused(norm(h), h)
        ref = ff.float()(torch.nn.functional.layer_norm(h.float(), (384,), norm.weight.float(), norm.bias.float(), norm.eps)) + h.float()
    assert rel_err(a, ref) < 2e-3 and rel_err(b, ref) < 2e-3


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
@pytest.mark.parametrize("m,k,n,act,res", [(2000, 256, 256, "none", True), (504, 384, 384, "none", True), (128, 640, 640, "none", True),
                                           (2000, 1024, 256, "none", True), (126, 2560, 640, "none", False), (4104, 768, 3072, "gelu", False),
                                           (513, 768, 3072, "gelu", False), (3000, 256, 512, "none", True), (300, 128, 512, "geglu", True),
                                           (1, 64, 64, "none", False), (129, 64, 128, "geglu", False)],
                         ids=["to_out256", "to_out384", "to_out640", "ff_out256", "ff_out640", "vit_fc1_b8", "vit_fc1_b1", "wide_res", "geglu_res",
                              "one_row", "two_tiles"])
def test_linear_16bit_epilogues_match_torch(m, k, n, act, res, dtype):
    """apx_linear with a 16-bit output (the row-per-thread TMA-store epilogue and the cluster split-K finisher of apx_gemm16.cu) against
    fp32 torch on the same 16-bit operands: bias, GELU / GEGLU, 16-bit residual with the caller's double rounding."""
    import torch.nn.functional as F
    from ap_adapter_b200 import _lib
    from gpu_util import record, rel_err
    g = torch.Generator().manual_seed(m * 7 + n)
    x = torch.randn(m, k, generator=g).to(dtype)
    w = (torch.randn(n, k, generator=g) * k ** -0.5).to(dtype)
    b = torch.randn(n, generator=g) * 0.2
    n_out = n // 2 if act == "geglu" else n
    r = torch.randn(m, n_out, generator=g).to(dtype) if res else None
    ref = F.linear(x.float(), w.float(), b)
    if act == "geglu":
        ref = ref[:, 0::2] * F.gelu(ref[:, 1::2])
    elif act == "gelu":
        ref = F.gelu(ref)
    if res:
        ref = ref.to(dtype).float() + r.float()
    n0 = _lib.launch_count()
    y = _lib.linear(x.cuda(), w.cuda(), b.cuda(), residual=r.cuda() if res else None, gelu=act == "gelu", geglu=act == "geglu")
    assert _lib.launch_count() - n0 == 1 and y.shape == (m, n_out)
    err = rel_err(y, ref)
    record(f"linear16/{m}x{k}x{n}/{act}", dtype=str(dtype), err=err)
    assert err < (1.5e-3 if dtype == torch.float16 else 1.2e-2)   # one or two roundings to the 16-bit output type


@pytest.mark.parametrize("m,k,n,geglu", [(504, 384, 3072, True), (128, 640, 5120, True), (2000, 256, 768, False)],
                         ids=["geglu_direct", "geglu_splitk", "qkv"])
def test_folded_layernorm_gemm_repeats_bit_for_bit(m, k, n, geglu):
    """The in-kernel row statistics read the TMA pipeline stages with plain ld.shared; releasing a stage while those loads were in
    flight gave wrong rows in ~1 of 5 launches of one instantiation (tools/stress_g16.py).  Every repeat must equal the first run
    bit for bit and sit within the tolerance of the fp32 reference, with other work in between."""
    import torch.nn.functional as F
    from ap_adapter_b200 import _lib
    from gpu_util import rel_err
    g = torch.Generator().manual_seed(m + n)
    x = (torch.randn(m, k, generator=g) * 1.5 + 0.8 * torch.randn(m, 1, generator=g) + 0.3).half()
    w = (torch.randn(n, k, generator=g) * k ** -0.5).half()
    b = torch.randn(n, generator=g) * 0.2
    gamma, beta = 1.0 + 0.3 * torch.randn(k, generator=g), 0.2 * torch.randn(k, generator=g)
    ref = F.linear(F.layer_norm(x.float(), (k,), gamma, beta, 1e-5), w.float(), b)
    if geglu:
        ref = ref[:, 0::2] * F.gelu(ref[:, 1::2])
    wg, c1, c2 = _lib.fold_layernorm(w.cuda(), b.cuda(), gamma.cuda(), beta.cuda())
    junk = torch.empty(32 << 20, device="cuda", dtype=torch.uint8)
    first = None
    for it in range(40):
        if it % 3 == 0:
            junk.random_(0, 255)
        y = _lib.linear_lnfold(x.cuda(), wg, c1, c2, 1e-5, geglu=geglu)
        if first is None:
            first = y
            assert rel_err(y, ref) < 2e-3
        else:
            assert torch.equal(y, first), f"run {it} differs from run 0"
