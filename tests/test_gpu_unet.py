"""The adapter processors inside the (restated) AudioLDM2 UNet on the GPU: one denoise step, the CUDA-graphed DDIM loop."""
import pytest
import torch

from ap_adapter_b200 import _lib
from ap_adapter_b200.attention_processor import AttnProcessor2_0, IPAttnProcessor2_0
from ap_adapter_b200.pipeline import APAdapterDenoiser
from ap_adapter_b200.unet import NHWC, AudioLDM2UNet2DConditionModel, BasicTransformerBlock, FeedForward, build_attn_processors
from oracle.ref_processors import OracleIPAttnProcessor2_0

from gpu_util import rel_err

pytestmark = pytest.mark.gpu


def _site_launches(bc=2, la=128) -> int:
    """libapx launches of the 32 adapter sites of one UNet forward at the 10 s latent: one per site where a fused kernel is picked
    (apx_ipattn_site_kernel != 0), three otherwise (SURVEY.md appendix B geometry: 10 x (1000, 256), 10 x (252, 384), 12 x (64, 640))."""
    lib = _lib.load()
    return sum(n * (1 if lib.apx_ipattn_site_kernel(bc, s, c, 8, 8, la) else 3) for n, s, c in ((10, 1000, 256), (10, 252, 384), (12, 64, 640)))


def _folded_norm_sites(bc=2) -> int:
    """Plain attention sites whose LayerNorm is folded into their Q(KV) projection (one extra libapx launch each): 14 plain sites per
    attention-bearing resnet layer (8 attn1 + 6 non-adapter attn2), at the levels with at least fold_norm_min_rows query rows."""
    n = AttnProcessor2_0.fold_norm_min_rows
    return 0 if n <= 0 else sum(14 * layers for layers, s in ((5, 1000), (5, 252), (6, 64)) if bc * s >= n)   # (adapter sites: no extra launch)


def _flash_sites(bc=2, heads=8) -> int:
    """Self-attention sites whose unmasked SDPA runs in libapx (apx_flash_attention): per attention-bearing resnet layer the 8 attn1
    sites and the 4 attn2 sites without a context (cross_attention_dim None), at the levels with at least flash_min_keys keys whose
    grid fits flash_max_ctas."""
    n, cap = AttnProcessor2_0.flash_min_keys, AttnProcessor2_0.flash_max_ctas
    dims = AttnProcessor2_0.flash_head_dims
    return sum(12 * layers for layers, s, d in ((5, 1000, 32), (5, 252, 48), (6, 64, 80)) if s >= n and d in dims and bc * heads * ((s + 127) // 128) <= cap)


def _add_ln_launches(bc=2) -> int:
    """apx_add_layernorm launches of one UNet forward: at the levels whose sites are below fold_norm_min_rows rows every transformer block
    (8 per attention-bearing resnet layer) runs norm1 as the LayerNorm-only form and attn1's residual add + norm2 as the fused form."""
    n = AttnProcessor2_0.fold_norm_min_rows
    if not BasicTransformerBlock.add_ln or n <= 0:
        return 0
    return sum(16 * layers for layers, s in ((5, 1000), (5, 252), (6, 64)) if bc * s < n)


def _fused_ff_count(unet) -> int:
    with torch.no_grad():
        return sum(m.fused_ok(torch.empty(1, m.net[2].out_features, device="cuda", dtype=torch.float16))
                   for m in unet.modules() if isinstance(m, FeedForward))


@pytest.fixture(scope="module")
def setup():
    torch.manual_seed(0)
    unet = AudioLDM2UNet2DConditionModel().eval().half().cuda()
    ours = build_attn_processors(unet, IPAttnProcessor2_0, ap_scale=0.5, device="cuda", dtype=torch.float16)
    theirs = build_attn_processors(unet, OracleIPAttnProcessor2_0, AttnProcessor2_0, ap_scale=0.5, device="cuda", dtype=torch.float16)
    for k, p in ours.items():
        if hasattr(p, "to_k_ip"):
            theirs[k].load_state_dict(p.state_dict())
    g = torch.Generator(device="cuda").manual_seed(1)
    inputs = dict(
        latents=torch.randn(1, 8, 250, 16, generator=g, device="cuda", dtype=torch.float16),
        gpe=torch.randn(2, 8 + 128, 768, generator=g, device="cuda", dtype=torch.float16),
        pe=torch.randn(2, 16, 1024, generator=g, device="cuda", dtype=torch.float16),
        mask=torch.ones(2, 16, device="cuda"))
    return unet, ours, theirs, inputs


def test_unet_step_matches_pytorch_processors(setup):
    unet, ours, theirs, inp = setup
    x = torch.cat([inp["latents"]] * 2)
    kw = dict(encoder_hidden_states=inp["gpe"], encoder_hidden_states_1=inp["pe"], encoder_attention_mask_1=inp["mask"])
    with torch.no_grad():
        unet.set_attn_processor(dict(theirs))
        # the comparison graph is the stock one: torch FeedForward, torch processors with separate to_q / to_k / to_v GEMMs
        FeedForward.use_fused = AttnProcessor2_0.fuse_projections = BasicTransformerBlock.fold_residual = NHWC.enabled = False
        add_ln, BasicTransformerBlock.add_ln = BasicTransformerBlock.add_ln, False
        try:
            y_ref = unet(x, torch.tensor(500, device="cuda"), **kw)[0]
        finally:
            FeedForward.use_fused = AttnProcessor2_0.fuse_projections = NHWC.enabled = True
            BasicTransformerBlock.fold_residual = "ip"
            BasicTransformerBlock.add_ln = add_ln
        unet.set_attn_processor(dict(ours))
        n0 = _lib.launch_count()
        y = unet(x, torch.tensor(500, device="cuda"), **kw)[0]
        launched = _lib.launch_count() - n0
    # every adapter site went through libapx: the C = 256 sites run the all-heads site kernel, the C = 384 sites the one-head
    # (cluster) site kernel, the C = 640 sites the three-kernel tcgen05 path; context caches / packed weights are built on first use
    n_ff = _fused_ff_count(unet)
    n_plain = sum(isinstance(p, AttnProcessor2_0) for p in ours.values())
    assert n_ff == 128 and n_plain == 224
    assert launched >= _site_launches() + 2 * n_ff + 109 + 22 + _folded_norm_sites() + _flash_sites() + _add_ln_launches(), launched   # + channels-last GroupNorm / output adds, folded norms
    assert torch.isfinite(y).all()
    # two fp16 evaluations of a 700 M-parameter network that differ only in the 32 adapter sites' rounding
    assert rel_err(y, y_ref) < 1e-2


def test_cuda_graph_loop_equals_eager_loop(setup):
    unet, ours, _, inp = setup
    unet.set_attn_processor(dict(ours))
    den = APAdapterDenoiser(unet)
    args = (inp["latents"], inp["gpe"], inp["pe"], inp["mask"])
    eager = den.denoise(*args, num_inference_steps=200, guidance_scale=7.5, use_cuda_graph=False, steps_to_run=3)
    n0 = _lib.launch_count()
    graphed = den.denoise(*args, num_inference_steps=200, guidance_scale=7.5, use_cuda_graph=True, steps_to_run=3)
    # same kernels, same order; cuBLAS / cuDNN may pick another algorithm under capture, so not asserted bitwise
    assert rel_err(graphed, eager) < 2e-3
    # replays launch kernels without passing through the C ABI again: only the 2 warm-up steps + the capture are counted
    per_step = den._graph["apx_launches_per_step"]
    # site kernel at the C = 256 sites, three kernels at C = 384 / 640, two GEMMs per GEGLU FeedForward (the plain sites' to_out +
    # residual GEMM is opt-in, APX_FOLD_RESIDUAL=1)
    n_ff = _fused_ff_count(unet)
    n_plain = sum(isinstance(p, AttnProcessor2_0) for p in ours.values()) if BasicTransformerBlock.fold_residual is True else 0
    # sites, FeedForwards, channels-last GroupNorms / output adds, Q(KV) projections with a folded LayerNorm, self-attention cores,
    # residual add + LayerNorm kernels of the small sites
    assert per_step == _site_launches() + 2 * n_ff + n_plain + 109 + 22 + _folded_norm_sites() + _flash_sites() + _add_ln_launches()
    assert _lib.launch_count() - n0 == 3 * per_step
    again = den.denoise(*args, num_inference_steps=200, guidance_scale=7.5, use_cuda_graph=True, steps_to_run=3)
    assert torch.equal(again, graphed)


def test_adapter_training_step_through_the_unet(setup):
    """BASELINE config 5 on one GPU: one bf16-autocast optimiser step through the restated UNet with the sm_100a processors,
    adapter-only gradients (train_apadapter_v2.py:892-979).  All 64 adapter tensors receive finite gradients, nothing else
    does, and repeating the step on the same batch lowers the loss."""
    from ap_adapter_b200 import training

    unet, ours, _, inp = setup
    unet.set_attn_processor(dict(ours))
    unet.float()                                   # fp32 weights, bf16 autocast compute: the reference's mixed-precision setup
    params = training.freeze_all_but_adapter(unet)
    assert len(params) == 64 and sum(p.numel() for p in params) == 21_626_880
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=1e-2)
    den = APAdapterDenoiser(unet)
    lat = inp["latents"].float()
    gpe, pe = inp["gpe"][:1].float(), inp["pe"][:1].float()
    mask = inp["mask"][:1]
    gen = torch.Generator(device="cuda")
    losses = []
    n0 = _lib.launch_count()
    for _ in range(3):
        gen.manual_seed(7)                         # same noise and timesteps every step: the loss must go down
        losses.append(training.train_step(unet, den.scheduler, opt, lat, gpe, pe, mask, generator=gen, params=params))
    assert _lib.launch_count() > n0
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in params)
    assert all(p.grad is None for n, p in unet.named_parameters() if "processor" not in n)
    assert losses[-1] < losses[0], losses
    unet.half()
    for p in unet.parameters():
        p.requires_grad_(False)
