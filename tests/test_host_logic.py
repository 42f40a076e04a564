"""Host-side logic that needs no GPU: processor registry of the restated UNet, checkpoint formats, the DDIM table,
clip sharding, and the reference's own processors running inside the restated host graph."""
import os

import numpy as np
import pytest
import torch

from ap_adapter_b200 import weights
from ap_adapter_b200.attention_processor import AttnProcessor2_0, IPAttnProcessor2_0
from ap_adapter_b200.pipeline import APAdapterDenoiser, DDIMScheduler, shard_range
from ap_adapter_b200.unet import AUDIOLDM2_LARGE_CONFIG, AudioLDM2UNet2DConditionModel, build_attn_processors
from oracle import ref_shims
from oracle.ref_processors import OracleAttnProcessor2_0, OracleIPAttnProcessor2_0

TINY = dict(block_out_channels=(32, 64, 96, 128), norm_num_groups=16)


@pytest.fixture(scope="module")
def tiny_unet():
    torch.manual_seed(0)
    return AudioLDM2UNet2DConditionModel(**TINY).eval()


def test_registry_has_256_processors_and_32_adapter_sites():
    with torch.device("meta"):
        unet = AudioLDM2UNet2DConditionModel()
    keys = list(unet.attn_processors)
    assert len(keys) == 256 and all(k.endswith(".processor") for k in keys)
    procs = build_attn_processors(unet, OracleIPAttnProcessor2_0, OracleAttnProcessor2_0, ap_scale=0.5)
    sites = {k: p for k, p in procs.items() if isinstance(p, OracleIPAttnProcessor2_0)}
    assert len(sites) == 32
    by_c = {}
    for k, p in sites.items():
        by_c[p.hidden_size] = by_c.get(p.hidden_size, 0) + 1
        assert ".attn2.processor" in k and k.split(".attentions.")[1].split(".")[0] in ("1", "5", "9")
    assert by_c == {256: 10, 384: 10, 640: 12}
    assert sum(p.numel() for s in sites.values() for p in s.parameters()) == 21_626_880


def test_set_attn_processor_contract(tiny_unet):
    procs = build_attn_processors(tiny_unet, OracleIPAttnProcessor2_0, OracleAttnProcessor2_0)
    with pytest.raises(ValueError, match="does not match"):
        tiny_unet.set_attn_processor({k: v for k, v in list(procs.items())[:5]})
    n = len(procs)
    tiny_unet.set_attn_processor(procs)
    assert len(procs) == 0  # the reference pops the entries (modeling_audioldm2.py:554-567)
    reg = tiny_unet.attn_processors
    assert len(reg) == n and sum(isinstance(p, OracleIPAttnProcessor2_0) for p in reg.values()) == 32
    # registered Module processors are parameters of the UNet (AttnProcsLayers / optimizer see them)
    assert any("processor.to_k_ip.weight" in k for k in tiny_unet.state_dict())
    tiny_unet.set_attn_processor(OracleAttnProcessor2_0())


def test_adapter_checkpoint_formats_roundtrip(tmp_path):
    torch.manual_seed(1)
    names = ["down_blocks.1.attentions.1.transformer_blocks.0.attn2.processor", "mid_block.attentions.1.transformer_blocks.1.attn2.processor"]
    procs = {n: IPAttnProcessor2_0(64, n, 768, num_tokens=8) for n in names}
    procs["x.attn1.processor"] = AttnProcessor2_0()
    sd = weights.adapter_state_dict(procs)
    assert sorted(sd) == sorted(f"{n}.to_{kv}_ip.weight" for n in names for kv in "kv")
    path = tmp_path / "pytorch_model.bin"
    torch.save({k: v.clone() for k, v in sd.items()}, path)
    fresh = {n: IPAttnProcessor2_0(64, n, 768, num_tokens=8) for n in names}
    assert weights.load_adapter_checkpoint(fresh, str(path), dtype=torch.float16) == 2
    for n in names:
        assert fresh[n].to_k_ip.weight.dtype == torch.float16
        assert torch.equal(fresh[n].to_k_ip.weight.float(), procs[n].to_k_ip.weight.half().float())
    with pytest.raises(KeyError):
        weights.load_adapter_state_dict({"other.attn2.processor": IPAttnProcessor2_0(64, "o", 768)}, sd)
    # copied_cross_attention/<site>_{k,v}.bin and do_copy=True
    d = tmp_path / "copied_cross_attention"
    k, v = torch.randn(64, 768).half(), torch.randn(64, 768).half()
    weights.save_copied_cross_attention(str(d), names[0], k, v)
    os.environ["APX_COPIED_CROSS_ATTENTION"] = str(d)
    try:
        p = IPAttnProcessor2_0(64, names[0], 768, num_tokens=8, do_copy=True)
    finally:
        del os.environ["APX_COPIED_CROSS_ATTENTION"]
    assert p.to_k_ip.weight.dtype == torch.float32 and p.to_k_ip.weight.requires_grad
    assert torch.equal(p.to_k_ip.weight.data, k.float()) and torch.equal(p.to_v_ip.weight.data, v.float())
    assert [n for n, _ in p.named_parameters()] == ["to_k_ip.weight", "to_v_ip.weight"]


@pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference not present")
def test_reads_the_shipped_copied_cross_attention_files():
    site = "up_blocks.0.attentions.9.transformer_blocks.1.attn2.processor"
    k, v = weights.load_copied_cross_attention(os.path.join(ref_shims.REF_ROOT, "copied_cross_attention"), site)
    assert k.shape == (640, 768) and v.shape == (640, 768) and k.dtype == torch.float16


def test_ddim_table_matches_stepwise_formula():
    s = DDIMScheduler()
    s.set_timesteps(200)
    ts = s.timesteps
    assert len(ts) == 200 and ts[0] == 996 and ts[-1] == 1 and (ts[:-1] - ts[1:] == 5).all()
    coefs = s.coefficients()
    g = torch.Generator().manual_seed(0)
    x, eps = torch.randn(2, 8, 10, 4, generator=g), torch.randn(2, 8, 10, 4, generator=g)
    for i in (0, 57, 199):
        a_t = s.alphas_cumprod[ts[i]]
        a_p = s.alphas_cumprod[ts[i] - 5] if ts[i] - 5 >= 0 else s.final_alpha_cumprod
        x0 = (x - (1 - a_t).sqrt() * eps) / a_t.sqrt()
        ref = a_p.sqrt() * x0 + (1 - a_p).sqrt() * eps
        assert torch.allclose(DDIMScheduler.step_with(coefs[i], eps, x), ref, atol=1e-5)
        assert torch.allclose(s.step(eps, ts[i], x), ref, atol=1e-5)
    # a perfect epsilon prediction recovers x0 at the end of the chain (size-independent property)
    x0 = torch.randn(1, 8, 10, 4, generator=g)
    noise = torch.randn(1, 8, 10, 4, generator=g)
    a0 = s.alphas_cumprod[ts[0]]
    xt = a0.sqrt() * x0 + (1 - a0).sqrt() * noise
    for i in range(200):
        xt = DDIMScheduler.step_with(coefs[i], noise, xt)
    a_end = s.final_alpha_cumprod
    assert torch.allclose(xt, a_end.sqrt() * x0 + (1 - a_end).sqrt() * noise, atol=1e-4)


def test_shard_range_partitions_every_count():
    for n in (1, 4, 7, 32, 33):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_unet_shapes_follow_the_reference_geometry(tiny_unet):
    seen = []
    tiny_unet.set_attn_processor(build_attn_processors(tiny_unet, OracleIPAttnProcessor2_0, OracleAttnProcessor2_0, ap_scale=0.5))

    def hook(mod, args, kwargs, out):
        if isinstance(mod.processor, OracleIPAttnProcessor2_0):
            seen.append((args[0].shape[1], args[0].shape[2], kwargs["encoder_hidden_states"].shape[-1]))

    hs = [m.register_forward_hook(hook, with_kwargs=True) for m in tiny_unet.modules() if hasattr(m, "set_processor")]
    x = torch.randn(2, 8, 250, 16)
    with torch.no_grad():
        y = tiny_unet(x, torch.tensor(981), encoder_hidden_states=torch.randn(2, 8 + 32, 768),
                      encoder_hidden_states_1=torch.randn(2, 6, 1024), encoder_attention_mask_1=torch.ones(2, 6))[0]
    for h in hs:
        h.remove()
    assert y.shape == x.shape and torch.isfinite(y).all()
    # 10 s latent 250x16 -> S = 1000 / 252 / 64 at the three adapter widths
    assert len(seen) == 32
    assert sorted(set(seen)) == [(64, 128, 768), (252, 96, 768), (1000, 64, 768)]


@pytest.mark.skipif(not ref_shims.reference_processors_available(), reason="neither /root/reference nor oracle/_ref present")
def test_reference_processors_drop_into_the_restated_unet(tiny_unet):
    """The reference's OWN IPAttnProcessor2_0 / AttnProcessor2_0, registered exactly as inference.py:22-59 does, run inside
    the restated host graph and agree with the oracle processors: the duck-typed Attention exposes what they read."""
    ap = ref_shims.load_attention_processor()
    torch.manual_seed(3)
    ref_procs = build_attn_processors(tiny_unet, ap.IPAttnProcessor2_0, ap.AttnProcessor2_0, ap_scale=0.6)
    ora_procs = build_attn_processors(tiny_unet, OracleIPAttnProcessor2_0, OracleAttnProcessor2_0, ap_scale=0.6)
    for k, p in ref_procs.items():
        if hasattr(p, "to_k_ip"):
            ora_procs[k].load_state_dict(p.state_dict())
    x = torch.randn(2, 8, 40, 16)
    kw = dict(encoder_hidden_states=torch.randn(2, 8 + 8, 768), encoder_hidden_states_1=torch.randn(2, 5, 1024),
              encoder_attention_mask_1=torch.tensor([[1, 1, 1, 0, 0], [1, 1, 1, 1, 1.0]]))
    with torch.no_grad():
        tiny_unet.set_attn_processor(ref_procs)
        y_ref = tiny_unet(x, 500, **kw)[0]
        tiny_unet.set_attn_processor(ora_procs)
        y_ora = tiny_unet(x, 500, **kw)[0]
    assert ((y_ref - y_ora).abs().max() / y_ref.abs().max()).item() < 1e-5


def test_denoiser_conditioning_block_layout(tiny_unet):
    """pipeline_audioldm2.py:935-949: rows [0,B) = [negative text | AudioMAE(zeros)], rows [B,2B) = [text | AudioMAE(mel)]."""
    class FakeMAE(torch.nn.Module):
        def forward(self, mel, time_pool=None, freq_pool=None):
            n = 512 // (time_pool * freq_pool)
            return [mel.mean().expand(mel.shape[0], n, 768).clone(), torch.ones(mel.shape[0], n)]

    den = APAdapterDenoiser(tiny_unet, FakeMAE())
    g = torch.randn(6, 8, 768)
    out = den.audio_condition(torch.full((1, 1024, 128), 2.0), g, 4, 4)
    assert out.shape == (6, 8 + 32, 768)
    assert torch.equal(out[:, :8], g)
    assert (out[:3, 8:] == 0).all() and (out[3:, 8:] == 2.0).all()
    lat = den.prepare_latents(3, 1000, torch.float32, "cpu", torch.Generator().manual_seed(0))
    assert lat.shape == (3, 8, 250, 16)


def test_kaldi_mel_banks_match_torchaudio():
    """The host-built mel filterbank that apx_fbank consumes restates torchaudio's ``get_mel_banks`` (the third-party routine
    behind the reference's extract_kaldi_fbank_feature, AudioMAE.py:368-377)."""
    from torchaudio.compliance import kaldi

    from ap_adapter_b200.audiomae import kaldi_mel_banks

    ref, _ = kaldi.get_mel_banks(128, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    ours = kaldi_mel_banks()
    assert ours.shape == (128, 257) and torch.all(ours[:, 256] == 0)
    assert (ours[:, :256] - ref).abs().max().item() < 5e-5


def test_feedforward_geglu_weight_interleave_is_the_same_function():
    """FeedForward._packed(): rows (hidden_i, gate_i) interleaved so that the GEMM epilogue's pairwise a * gelu(g) over adjacent
    columns equals diffusers' GEGLU chunk(2) (the arithmetic of APX_EPI_GEGLU restated with torch ops on the packed weight)."""
    import torch.nn.functional as F
    from ap_adapter_b200.unet import FeedForward
    torch.manual_seed(0)
    ff = FeedForward(64).eval()
    x = torch.randn(3, 10, 64)
    w1, b1, b2 = ff._packed()
    assert w1.shape == (512, 64) and b1.shape == (512,) and b1.dtype == torch.float32 and b2.shape == (64,)
    z = F.linear(x, w1, b1)
    hidden = z[..., 0::2] * F.gelu(z[..., 1::2])
    got = F.linear(hidden, ff.net[2].weight, b2) + x
    with torch.no_grad():
        want = ff(x) + x
    assert torch.allclose(got, want, atol=1e-5)
    # CPU tensors / grad mode never take the libapx route; the cache follows in-place weight updates
    assert not ff.fused_ok(x)
    packed_before = ff._packed()[0]
    with torch.no_grad():
        ff.net[0].proj.weight.mul_(2.0)
    assert ff._packed()[0] is not packed_before and torch.equal(ff._packed()[0][0::2], ff.net[0].proj.weight[:256])


def test_plain_processor_projection_fusion_and_residual_argument_on_cpu():
    """AttnProcessor2_0: the concatenated QKV / KV weight is rebuilt when a projection changes, and `residual_inout` on the stock
    (CPU / grad) route is a plain add (reference attention_processor.py:199-294 plus the caller's residual add)."""
    from ap_adapter_b200.attention_processor import AttnProcessor2_0
    from ap_adapter_b200.unet import Attention
    torch.manual_seed(1)
    attn = Attention(64, None, 4, 16).eval()
    proc = AttnProcessor2_0()
    attn.set_processor(proc)
    w = proc._cat_weight((attn.to_q, attn.to_k, attn.to_v))
    assert w.shape == (192, 64) and torch.equal(w[64:128], attn.to_k.weight)
    assert proc._cat_weight((attn.to_q, attn.to_k, attn.to_v)) is w
    with torch.no_grad():
        attn.to_k.weight.add_(1.0)
    w2 = proc._cat_weight((attn.to_q, attn.to_k, attn.to_v))
    assert w2 is not w and torch.equal(w2[64:128], attn.to_k.weight)
    x, h = torch.randn(2, 9, 64), torch.randn(2, 9, 64)
    with torch.no_grad():
        y = attn(x)
        q, k, v = proc._project(attn, x, x, True)          # CPU: three separate projections
        assert torch.equal(q, attn.to_q(x)) and torch.equal(k, attn.to_k(x)) and torch.equal(v, attn.to_v(x))
        assert torch.equal(attn(x, residual_inout=h), y + h)


def test_bench_torch_gpu_comparison_processor_is_the_reference_algorithm():
    """bench.py's `torch_gpu` leg times a bench-local PyTorch restatement of the adapter processor; it must compute what the
    oracle (pinned against the reference's own outputs) computes."""
    import bench
    from ap_adapter_b200.unet import Attention
    from oracle.ref_processors import OracleIPAttnProcessor2_0
    torch.manual_seed(2)
    attn = Attention(64, 768, 4, 16).eval()
    a = bench._TorchIPAttnSite(hidden_size=64, name="s", cross_attention_dim=768, num_tokens=8, scale=0.5)
    b = OracleIPAttnProcessor2_0(hidden_size=64, name="s", cross_attention_dim=768, num_tokens=8, scale=0.5)
    b.load_state_dict(a.state_dict())
    x, ctx = torch.randn(2, 33, 64), torch.randn(2, 8 + 24, 768)
    with torch.no_grad():
        assert torch.allclose(a(attn, x, encoder_hidden_states=ctx), b(attn, x, encoder_hidden_states=ctx), atol=1e-5)


def test_audiomae_checkpoint_loader_reads_the_reference_format(tmp_path, monkeypatch):
    """``pretrained.pth`` = ``{'model': state_dict}`` with timm key names, loaded non-strict from the working directory
    (audio_encoder/AudioMAE.py:28-30): encoder keys land in the parameter tree, decoder / mask-token keys are ignored."""
    from ap_adapter_b200.audiomae import AudioMAEConditionCTPoolRand, Vanilla_AudioMAE
    from ap_adapter_b200.weights import load_audiomae_checkpoint
    from oracle import cases

    sd = cases.mae_state_dict(depth=2)
    full = dict(sd)
    full["mask_token"] = torch.zeros(1, 1, 512)
    full["decoder_embed.weight"] = torch.randn(512, 768)
    full["decoder_blocks.0.attn.qkv.weight"] = torch.randn(1536, 512)
    path = tmp_path / "pretrained.pth"
    torch.save({"model": full, "epoch": 3}, path)
    loaded = load_audiomae_checkpoint(str(path))
    assert set(loaded) == set(sd) and all(torch.equal(loaded[k], sd[k]) for k in sd)
    m = Vanilla_AudioMAE(checkpoint_path=str(path), depth=2)
    got = m.model.state_dict()
    assert set(got) == set(sd)
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    # the reference's default: a bare "pretrained.pth" resolved against the current directory
    monkeypatch.chdir(tmp_path)
    m2 = AudioMAEConditionCTPoolRand(depth=2)
    assert torch.equal(m2.audiomae.model.blocks[1].mlp.fc2.weight, sd["blocks.1.mlp.fc2.weight"])
    assert all(not p.requires_grad for p in m2.audiomae.parameters())
    # a bare state dict (no 'model' wrapper) is accepted too
    torch.save(full, tmp_path / "bare.pth")
    assert set(load_audiomae_checkpoint(str(tmp_path / "bare.pth"))) == set(sd)
    with pytest.raises(FileNotFoundError):
        Vanilla_AudioMAE(checkpoint_path=str(tmp_path / "missing.pth"), depth=1)


def test_step_graph_key_covers_what_a_captured_step_bakes_in(tiny_unet):
    """APAdapterDenoiser._site_state: ap_scale, num_tokens and the adapter weights' identity / version are part of the graph key."""
    from ap_adapter_b200.attention_processor import IPAttnProcessor2_0
    procs = build_attn_processors(tiny_unet, IPAttnProcessor2_0, ap_scale=0.5)
    tiny_unet.set_attn_processor(dict(procs))
    den = APAdapterDenoiser(tiny_unet)
    adapters = den._adapter_procs()
    assert len(adapters) == sum(hasattr(p, "to_k_ip") for p in procs.values()) > 0
    s0 = den._site_state()
    assert den._site_state() == s0
    adapters[0].scale = 0.75
    s1 = den._site_state()
    assert s1 != s0
    with torch.no_grad():
        adapters[-1].to_k_ip.weight.add_(1.0)
    s2 = den._site_state()
    assert s2 != s1
    adapters[1].to_v_ip.weight = torch.nn.Parameter(torch.zeros_like(adapters[1].to_v_ip.weight))
    assert den._site_state() != s2
    assert den._site_blobs() == [None] * (3 * len(adapters))
    den.invalidate_graph()
    assert den._graph is None


def test_channels_last_route_is_the_same_function_as_the_stock_graph(tiny_unet, monkeypatch):
    """The channels-last inference route of the restated UNet (NHWC: fused GroupNorm(+bias+temb)+SiLU, batched time-embedding
    GEMM, proj_in / proj_out as linears on the free token view, shortcut + conv2 + biases in one add) computes what the stock NCHW
    graph computes.  Run here on the CPU in fp32 with the two libapx kernels replaced by their torch statements, so the WIRING
    (which bias goes where, slice offsets of the batched GEMM, views, memory formats) is what is tested; the kernels themselves
    are pinned on the GPU (tests/test_gpu_unet.py)."""
    from ap_adapter_b200 import _lib, unet as U

    def add_bias(a, b, bias, scale=1.0):
        y = a + b if bias is None else a + b + bias.to(a.dtype)[None, :, None, None]
        return (y * scale).contiguous(memory_format=torch.channels_last)

    monkeypatch.setattr(U.NHWC, "usable", staticmethod(lambda x: U.NHWC.enabled and not torch.is_grad_enabled()))
    monkeypatch.setattr(_lib, "groupnorm_nhwc_ok", lambda hw, c, g: False)      # -> the torch statement inside _gn_nhwc
    monkeypatch.setattr(_lib, "add_bias_nhwc", add_bias)
    tiny_unet.set_attn_processor(build_attn_processors(tiny_unet, OracleIPAttnProcessor2_0, OracleAttnProcessor2_0, ap_scale=0.5))
    torch.manual_seed(4)
    for p in tiny_unet.parameters():            # non-trivial biases / affine parameters everywhere
        if p.dim() == 1:
            p.data.normal_(0.0, 0.3)
    x = torch.randn(2, 8, 50, 16)               # 50 rows: the up path needs forward_upsample_size
    kw = dict(encoder_hidden_states=torch.randn(2, 8 + 8, 768), encoder_hidden_states_1=torch.randn(2, 5, 1024),
              encoder_attention_mask_1=torch.tensor([[1, 1, 1, 0, 0], [1, 1, 1, 1, 1.0]]))
    seen = []
    hook = tiny_unet.mid_block.resnets[0].register_forward_hook(lambda m, i, o: seen.append(o.is_contiguous(memory_format=torch.channels_last)))
    with torch.no_grad():
        monkeypatch.setattr(U.NHWC, "enabled", False)
        y_stock = tiny_unet(x, 321, **kw)[0]
        monkeypatch.setattr(U.NHWC, "enabled", True)
        y_nhwc = tiny_unet(x, 321, **kw)[0]
    hook.remove()
    assert seen == [False, True]                # the second forward really ran channels-last
    assert y_nhwc.is_contiguous() and y_nhwc.shape == y_stock.shape
    assert ((y_nhwc - y_stock).abs().max() / y_stock.abs().max()).item() < 1e-5
    assert not U.NHWC.active and all("_temb_proj" not in r.__dict__ for r in tiny_unet._resnets())


def test_sinc_resample_taps_match_torchaudio():
    """The host-built polyphase table apx_resample consumes restates torchaudio's ``_get_sinc_resample_kernel`` (the third-party
    routine behind the reference's resample call, AudioMAE.py:360-366)."""
    import math

    import torchaudio.functional as AF

    from ap_adapter_b200.audiomae import sinc_resample_taps
    for sr in (48000, 44100, 22050, 8000, 32000):
        g = math.gcd(sr, 16000)
        ref, width = AF.functional._get_sinc_resample_kernel(sr, 16000, g)
        taps, orig, new, w = sinc_resample_taps(sr, 16000)
        assert (orig, new, w) == (sr // g, 16000 // g, width) and taps.shape == ref[:, 0, :].shape
        assert (taps - ref[:, 0, :]).abs().max().item() < 1e-5


def test_invalidate_caches_drops_every_derived_weight_copy(tiny_unet):
    """The derived tensors (packed GEGLU weights, concatenated QKV, fp32 copies, additive masks) are keyed on identity + ``_version``;
    a ``param.data`` write (8-bit / fused optimisers, EMA copies: train_apadapter_v2.py:733-741) is invisible to the key, and
    ``AudioLDM2UNet2DConditionModel.invalidate_caches`` / ``APAdapterDenoiser.invalidate_graph`` is the documented way to drop them."""
    import ap_adapter_b200.unet as U
    from ap_adapter_b200.attention_processor import IPAttnProcessor2_0
    tiny_unet.set_attn_processor(build_attn_processors(tiny_unet, IPAttnProcessor2_0, ap_scale=0.5))
    ff = next(m for m in tiny_unet.modules() if isinstance(m, U.FeedForward))
    attn = next(m for m in tiny_unet.modules() if hasattr(m, "set_processor") and isinstance(m.processor, AttnProcessor2_0))
    plain = attn.processor
    conv = tiny_unet.conv_in
    mask = torch.ones(2, 16)

    w_ff = ff._packed()[0]
    w_cat = plain._cat_weight((attn.to_q, attn.to_k, attn.to_v))
    b32 = U._f32(conv, "bias")
    with torch.no_grad():
        m0 = tiny_unet._additive_mask(mask, torch.float32, "m0")
    # versioned in-place edits are seen by the keys ...
    with torch.no_grad():
        ff.net[0].proj.weight.mul_(2.0)
    assert ff._packed()[0] is not w_ff
    w_ff = ff._packed()[0]
    # ... `.data` writes are not: the stale copies are served until the caches are dropped
    ff.net[0].proj.weight.data.mul_(0.5)
    attn.to_k.weight.data.add_(1.0)
    conv.bias.data.add_(1.0)
    assert ff._packed()[0] is w_ff and plain._cat_weight((attn.to_q, attn.to_k, attn.to_v)) is w_cat and U._f32(conv, "bias") is b32
    den = APAdapterDenoiser(tiny_unet)
    den._graph = dict(key=None)
    den.invalidate_graph()                     # a new prompt: the graph and the adapter sites' caches only
    assert den._graph is None and ff._packed()[0] is w_ff
    den._graph = dict(key=None)
    den.invalidate_graph(unet_weights=True)    # `.data` edits of the UNet: every derived copy
    assert den._graph is None and plain._cat is None and ff._fused is None
    assert all(U._CACHE_ATTR not in m.__dict__ for m in tiny_unet.modules())
    inner = ff.net[0].proj.weight.shape[0] // 2
    assert torch.equal(ff._packed()[0][0::2], ff.net[0].proj.weight[:inner])
    assert torch.equal(plain._cat_weight((attn.to_q, attn.to_k, attn.to_v))[inner_rows(attn):2 * inner_rows(attn)], attn.to_k.weight)
    assert torch.equal(U._f32(conv, "bias"), conv.bias.float())
    with torch.no_grad():
        assert tiny_unet._additive_mask(mask, torch.float32, "m0") is not m0
    for p in tiny_unet.attn_processors.values():
        if isinstance(p, IPAttnProcessor2_0):
            assert p._ctx_cache is None and p._w_cache is None and p._wln_cache is None and p._bo_cache is None
    attn.to_k.weight.data.sub_(1.0)   # the fixture is shared by the module
    conv.bias.data.sub_(1.0)
    tiny_unet.invalidate_caches()


def inner_rows(attn):
    return attn.to_q.weight.shape[0]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_layernorm_fold_constants_reproduce_layernorm_then_linear(dtype):
    """``_lib.fold_layernorm`` is host arithmetic: with Wg = W * gamma (rounded to the kernel dtype), c1 = rowsum(Wg), c2 = W beta + b the
    GEMM epilogue's ``rstd * (x Wg^T - mean * c1) + c2`` is LayerNorm(x) W^T + b (norm1 / norm2 / norm3 of BasicTransformerBlock in front
    of to_q / to_k / to_v and the GEGLU projection: modeling_audioldm2 call sites 1140-1156).  Checked in fp64 on the CPU against
    torch's own layer_norm + linear, with the rounded Wg on both sides so that only the algebra is tested."""
    from ap_adapter_b200 import _lib
    g = torch.Generator().manual_seed(3)
    m, k, n, eps = 37, 320, 96, 1e-5
    x = (torch.randn(m, k, generator=g) * 1.5 + 0.7).to(dtype)
    w = (torch.randn(n, k, generator=g) * k ** -0.5).to(dtype)
    b = torch.randn(n, generator=g) * 0.2
    gamma, beta = 1.0 + 0.3 * torch.randn(k, generator=g), 0.2 * torch.randn(k, generator=g)
    wg, c1, c2 = _lib.fold_layernorm(w, b, gamma, beta)
    assert wg.dtype == dtype and c1.dtype == torch.float32 and c2.dtype == torch.float32 and c1.shape == (n,) and c2.shape == (n,)
    xd = x.double()
    mean, var = xd.mean(dim=1, keepdim=True), xd.var(dim=1, unbiased=False, keepdim=True)
    folded = (var + eps).rsqrt() * (xd @ wg.double().t() - mean * c1.double()) + c2.double()
    # the same function with the rounded, gamma-scaled weight un-folded: LN without affine, then Wg, then the beta / bias term
    ref = torch.nn.functional.layer_norm(xd, (k,), None, None, eps) @ wg.double().t() + (w.double() @ beta.double() + b.double())
    assert ((folded - ref).abs().max() / ref.abs().max()).item() < 1e-6
    # and against the unrounded module chain it stays within the rounding of Wg
    full = torch.nn.functional.linear(torch.nn.functional.layer_norm(xd, (k,), gamma.double(), beta.double(), eps), w.double(), b.double())
    assert ((folded - full).abs().max() / full.abs().max()).item() < (2e-3 if dtype == torch.float16 else 1.5e-2)
