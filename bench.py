#!/usr/bin/env python
"""bench.py — UNet denoise steps/s (10 s clip, 200-step DDIM schedule) with the AP-adapter hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one denoise step of BASELINE.json configs[1]: CFG concat -> AudioLDM2-large UNet (random init, the geometry
the shipped adapter weights imply, fp16) with the 32 adapter sites running in libapx.so -> guidance -> DDIM update, for
B clips per GPU (default 1) of 10 s (latent 8x250x16), text context 8 tokens + 128 AudioMAE tokens (timbre_transfer
preset: time/freq pooling 2), T5 context 16 tokens.  Ranks run independent clips (weak scaling); the only collective
is the final gather of latents.

Prints ONE JSON line (rank 0).  `value` is timed on the device with inputs resident in HBM; `e2e` runs the same steps
through the public API with the step's latents coming from / going back to pinned host memory every step.
`--impl reference` times the reference algorithm (oracle port: same UNet, fp32, reference-semantics processors) on the
host CPUs.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "unet_denoise_steps_per_s"
UNIT = "steps/s"
LT, LA, L_T5 = 8, 128, 16
LAT_SHAPE = (8, 250, 16)
INFER_STEPS = 200
GUIDANCE = 7.5
AP_SCALE = 0.5
WORKLOAD = "step"
WORKLOAD_NAME = ("BASELINE configs[1]: AudioLDM2-large UNet single denoise step (CFG, DDIM-200 schedule) with AP "
                 "attn2 processors at 32 sites, 10 s latent (8x250x16)")


def set_workload(name: str, args) -> None:
    """`step` (default) = BASELINE configs[1], the configuration the metric is quoted on.  The other presets re-parameterise the same
    step loop (config4) or time a different unit of work (config3: one whole generation; train: one optimiser step); they are run
    by hand / by tools/gpu_visit.sh and their lines are committed under profiles/, the driver's round-end line stays `step`."""
    global LA, AP_SCALE, GUIDANCE, WORKLOAD, WORKLOAD_NAME
    WORKLOAD = name
    if name == "config4":      # style_transfer preset (reference config.py:30-33): pooling 4x4 -> 32 audio tokens, ap_scale 0.55, g 9.5
        LA, AP_SCALE, GUIDANCE = 32, 0.55, 9.5
        if args.batch == 1:
            args.batch = 4     # 32 clips over 8 GPUs
        WORKLOAD_NAME = ("BASELINE configs[3]: style_transfer preset, batch sharded data-parallel (4 clips per GPU), UNet denoise step with AP "
                         "attn2 processors, 10 s latent (8x250x16); NCCL gather of the final latents only")
    elif name == "config3":
        if args.batch == 1:
            args.batch = 4
        WORKLOAD_NAME = ("BASELINE configs[2]: full timbre_transfer inference, AudioMAE encode (mel + zero mel) + context K/V caches + 200 "
                         "CUDA-graphed CFG/DDIM steps, batch 4, through APAdapterDenoiser.__call__ (host mel in, host latents out)")
    elif name == "train":
        if args.batch == 1:
            args.batch = 4
        WORKLOAD_NAME = ("BASELINE configs[4]: one adapter-training optimiser step (train_apadapter_v2.py:892-979), bf16 autocast, fp32 adapter "
                         "weights, adapter-only gradients, 4 clips per GPU, bucketed NCCL all-reduce overlapped with the backward")


def workload_config(batch: int, n_gpus: int) -> dict:
    return {"workload": WORKLOAD_NAME,
            "clips_per_gpu": batch, "cfg_batch_per_gpu": 2 * batch, "text_tokens": LT, "audio_tokens": LA,
            "t5_tokens": L_T5, "ap_scale": AP_SCALE, "guidance_scale": GUIDANCE, "ddim_steps": INFER_STEPS,
            "parallelism": f"clips sharded over {n_gpus} rank(s), no data-path collective, final all-gather of latents",
            "l2": "per-step working set (1.4 GB of fp16 UNet weights) exceeds the 126 MB L2; no flush needed"}


HOST_GRAPH = ("adapter sites: libapx (all-heads site kernel at C=256, one-head cluster site kernel at C=384, 3-kernel form at C=640; residual "
              "add folded in); GEGLU FeedForwards: 2 libapx GEMMs with GEGLU / residual epilogues; convolutional trunk channels-last with "
              "libapx GroupNorm(+conv bias + time embedding)+SiLU and ResnetBlock output-add kernels, proj_in / proj_out as linears on the "
              "token view, one GEMM for all time-embedding projections; LayerNorms folded into the libapx GEMMs (norm3 -> GEGLU projection "
              "everywhere, norm1 / norm2 -> Q(KV) projection at sites with >= 1000 rows, norm2 -> site kernel at C=256); plain attention "
              "sites: fused QKV projection (cuBLAS below 1000 rows), self-attention core in libapx (apx_flash_attention) while its grid fits "
              "one wave, cuDNN SDPA otherwise and at the masked T5 sites, to_out on cuBLAS; at the sites below 1000 rows attn1's residual add + norm2 in one libapx kernel and norm1 as its LayerNorm-only form; T5 K/V projected once per context; one CUDA "
              "graph per step")


def peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d["bf16_tflops"], "tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(batch: int, steps: int, warmup: int, budget_s: float) -> dict:
    """The reference algorithm on the host CPUs: restated UNet in fp32 with the oracle port of IPAttnProcessor2_0 /
    AttnProcessor2_0 at every site (the reference itself is pure Python on diffusers, which this image lacks)."""
    from ap_adapter_b200.pipeline import APAdapterDenoiser
    from ap_adapter_b200.unet import AudioLDM2UNet2DConditionModel, build_attn_processors
    from oracle import ref_shims
    from oracle.ref_processors import OracleAttnProcessor2_0, OracleIPAttnProcessor2_0

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    unet = AudioLDM2UNet2DConditionModel().eval()
    if ref_shims.reference_processors_available():
        # oracle/_ref: the reference's OWN IPAttnProcessor2_0 / AttnProcessor2_0 classes (its attention_processor module, compiled by oracle/build_ref.py) at the
        # 256 sites; only the UNet around them is the restated host graph (diffusers is absent from this image)
        ap = ref_shims.load_attention_processor()
        ip_cls, plain_cls, procs = ap.IPAttnProcessor2_0, ap.AttnProcessor2_0, "reference classes (oracle/_ref)"
    else:
        ip_cls, plain_cls, procs = OracleIPAttnProcessor2_0, OracleAttnProcessor2_0, "oracle port (oracle/ref_processors.py)"
    unet.set_attn_processor(build_attn_processors(unet, ip_cls, plain_cls, ap_scale=AP_SCALE))
    den = APAdapterDenoiser(unet)
    g = torch.Generator().manual_seed(1)
    lat = torch.randn(batch, *LAT_SHAPE, generator=g)
    gpe = torch.randn(2 * batch, LT + LA, 768, generator=g)
    pe = torch.randn(2 * batch, L_T5, 1024, generator=g)
    mask = torch.ones(2 * batch, L_T5)
    den.scheduler.set_timesteps(INFER_STEPS)
    ts, coefs = den.scheduler.timesteps, den.scheduler.coefficients()
    times = []
    with torch.no_grad():
        i = 0
        for _ in range(max(1, warmup)):
            t0 = time.perf_counter()
            lat = den.denoise_step(lat, ts[i % INFER_STEPS], coefs[i % INFER_STEPS], gpe, pe, mask, GUIDANCE)
            est = time.perf_counter() - t0
            i += 1
            if est * (warmup + steps) > budget_s:
                break  # slow host: one warm-up step is all the budget allows
        n = max(1, min(steps, int(budget_s / max(est, 1e-3))))
        for _ in range(n):
            t0 = time.perf_counter()
            lat = den.denoise_step(lat, ts[i % INFER_STEPS], coefs[i % INFER_STEPS], gpe, pe, mask, GUIDANCE)
            times.append(time.perf_counter() - t0)
            i += 1
    assert torch.isfinite(lat).all()
    total = sum(times)
    return {"value": batch * len(times) / total, "unit": UNIT, "cores": cores, "kind": "port", "processors": procs,
            "sample": f"{len(times)} full fp32 UNet denoise steps of the same workload (B={batch}) after {i - len(times)} warm-up, "
                      f"torch {torch.__version__} with {torch.get_num_threads()} threads",
            "ms_per_step": 1e3 * total / len(times), "steps_timed": len(times)}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.rows = []
        self.proc = None
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            ident = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
        except Exception:
            ident = str(device_index)
        # 200 ms period: at 50 ms the polling itself cost the device-timed loop ~4 % (13.9 vs 13.3 ms per step with / without it)
        self.cmd = ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200", "-i", ident]

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(self.cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, t0: float, t1: float) -> dict:
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 6] or [r for _, r in self.rows if len(r) >= 6]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(rows[0][1]) if rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(rows)}


# ------------------------------------------------------------------------------------------------ our arm
def gemm_roofline(batch: int, iters: int = 30) -> dict:
    """The kernel with the largest share of the step's device time after the site kernels: libapx's GEMM, at its heaviest call of
    the step — the C = 256 FeedForward projection with norm3 folded in and the GEGLU epilogue (40 calls per step).  Same method as
    site_roofline: CUDA events around replays of a graph that rotates over operand copies larger than the L2."""
    from ap_adapter_b200 import _lib

    dev = torch.device("cuda", torch.cuda.current_device())
    m, n, k = 2 * batch * 1000, 2048, 256
    per_copy = 2 * (m * k + n * k) + 2 * m * (n // 2)
    r = max(4, min(256, int(2.5 * 126e6 / per_copy) + 1))
    g = torch.Generator(device=dev).manual_seed(1)
    a = torch.randn(r, m, k, generator=g, device=dev, dtype=torch.float16)
    w = torch.randn(r, n, k, generator=g, device=dev, dtype=torch.float16) * k ** -0.5
    bias = torch.zeros(n, device=dev, dtype=torch.float16)
    gamma, beta = torch.ones(k, device=dev, dtype=torch.float16), torch.zeros(k, device=dev, dtype=torch.float16)
    folded = [_lib.fold_layernorm(w[i], bias, gamma, beta) for i in range(r)]

    def launch(i):
        wg, c1, c2 = folded[i]
        return _lib.linear_lnfold(a[i], wg, c1, c2, 1e-5, geglu=True)

    for i in range(r):
        launch(i)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs = [launch(i) for i in range(r)]   # kept alive: the graph's private pool owns them
    for _ in range(3):
        graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    del outs
    us = 1e3 * e0.elapsed_time(e1) / (iters * r)
    flops = 2.0 * m * n * k
    byts = per_copy + 8 * n
    pk = peaks()
    t_tensor, t_hbm = flops / (pk["tflops"] * 1e12), byts / (pk["hbm_gbs"] * 1e9)
    bound = "tensor" if t_tensor >= t_hbm else "hbm"
    if bound == "tensor":
        achieved, peak, unit = flops / (us * 1e-6) / 1e12, pk["tflops"], "TFLOP/s"
    else:
        achieved, peak, unit = byts / (us * 1e-6) / 1e9, pk["hbm_gbs"], "GB/s"
    return {"kernel": f"gemm16_kernel<128, GEGLU, LN> via apx_linear_lnfold: LayerNorm + Linear + GEGLU of the C=256 FeedForward, M={m} N={n} K={k} fp16",
            "bound": bound, "achieved": round(achieved, 3), "peak": peak, "unit": unit, "frac": round(achieved / peak, 4), "traffic": None,
            "us_per_call": round(us, 2), "algorithmic_bytes": byts, "algorithmic_flops": flops,
            "floor_us": round(max(t_tensor, t_hbm) * 1e6, 3), "peak_source": pk["source"], "operand_copies_rotated": r,
            "calls_per_step": 40}


def site_roofline(batch: int, iters: int = 30) -> dict:
    """The dominant kernel of the hot path: the adapter site at the S=1000, C=256 geometry (10 of the 32 sites, and the
    one that moves the most bytes).  Timed through apx_ipattn_fwd with CUDA events around replays of a CUDA graph that
    rotates over enough distinct copies of every operand (X, Y, Wq, Wo, context cache, workspace) that no launch finds
    its inputs in the 126 MB L2."""
    from ap_adapter_b200 import _lib

    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    bc, s, c, h, d = 2 * batch, 1000, 256, 8, 768
    per_copy = 4 * bc * s * c * 2 + 2 * c * c * 2
    r = max(4, min(256, int(2.5 * 126e6 / per_copy) + 1))
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(r, bc, s, c, generator=g, device=dev, dtype=torch.float16)
    y = torch.empty_like(x)
    wq = torch.randn(r, c, c, generator=g, device=dev, dtype=torch.float16) * c ** -0.5
    wo = torch.randn(r, c, c, generator=g, device=dev, dtype=torch.float16) * c ** -0.5
    bo = torch.zeros(c, device=dev)
    ctx = torch.randn(bc, LT + LA, d, generator=g, device=dev, dtype=torch.float16)
    wkv = [torch.randn(c, d, generator=g, device=dev, dtype=torch.float16) * d ** -0.5 for _ in range(4)]
    cbytes = lib.apx_ctx_kv_bytes(bc, LT, LA, c, h)
    cache = torch.empty(r, (cbytes + 255) // 256 * 256, dtype=torch.uint8, device=dev)
    wsb = max(lib.apx_ctx_kv_workspace_bytes(bc, LT, LA, c), lib.apx_ipattn_workspace_bytes(bc, s, c))
    ws = torch.empty(r, (wsb + 255) // 256 * 256, dtype=torch.uint8, device=dev)
    st = _lib.stream_ptr()
    pbytes = lib.apx_ipattn_pack_bytes(c, h)
    fused = bool(pbytes) and lib.apx_ipattn_fused_ok(c, h, LT, LA) == 1
    pack = torch.empty(r, max(256, (pbytes + 255) // 256 * 256), dtype=torch.uint8, device=dev)
    for i in range(r):
        if fused:
            _lib.check(lib.apx_ipattn_pack(_lib.ptr(wq[i]), _lib.ptr(wo[i]), _lib.ptr(pack[i]), c, h, 0, st), "apx_ipattn_pack")
        _lib.check(lib.apx_ctx_kv(_lib.ptr(ctx), *[_lib.ptr(w) for w in wkv], _lib.ptr(cache[i]), bc, LT, LA, d, c, h, 0,
                                  _lib.ptr(ws[i]), wsb, st), "apx_ctx_kv")

    def launch(i):
        _lib.check(lib.apx_ipattn_fwd(_lib.ptr(x[i]), _lib.ptr(wq[i]), _lib.ptr(cache[i]), _lib.ptr(wo[i]),
                                      _lib.ptr(pack[i]) if fused else None, _lib.ptr(bo), _lib.ptr(y[i]), bc, s, c, h, LT, LA,
                                      AP_SCALE, 0, _lib.ptr(ws[i]), wsb, _lib.stream_ptr()), "apx_ipattn_fwd")   # the CURRENT stream: capture

    for i in range(r):
        launch(i)
    torch.cuda.synchronize()
    # One CUDA graph holds r launches (one per operand copy); replaying it removes the host-side submission cost (ctypes call,
    # tensor-map encoding), which would otherwise bound the measurement of a microsecond-scale kernel.
    n0 = _lib.launch_count()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(r):
            launch(i)
    launches_per_call = (_lib.launch_count() - n0) // r
    for _ in range(3):
        graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / (iters * r)
    l_tot = LT + LA
    flops = 4.0 * bc * s * c * c + 4.0 * bc * s * l_tot * c                # q-proj + out-proj + QK^T + PV
    l_pad = (LT + 15) // 16 * 16 + (LA + 15) // 16 * 16
    kv_bytes = 2 * bc * h * l_pad * (c // h) * 2                             # packed K + V^T tiles as read
    byts = 2 * bc * s * c * 2 + 2 * c * c * 2 + kv_bytes + 4 * c           # X in, Y out, Wq, Wo, K/V cache, bias
    pk = peaks()
    t_tensor, t_hbm = flops / (pk["tflops"] * 1e12), byts / (pk["hbm_gbs"] * 1e9)
    bound = "tensor" if t_tensor >= t_hbm else "hbm"
    if bound == "tensor":
        achieved, peak, unit = flops / (us * 1e-6) / 1e12, pk["tflops"], "TFLOP/s"
    else:
        achieved, peak, unit = byts / (us * 1e-6) / 1e9, pk["hbm_gbs"], "GB/s"
    name = "ipattn_site_kernel (to_q + dual softmax(QK^T)V + ap_scale add + to_out in one tcgen05 kernel)" if fused else \
        "gemm_tc + ipattn_core + gemm_tc"
    # DRAM traffic of the same launch from the committed ncu capture (null when this shape was not captured)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = tj.get(f"ipattn_site_kernel Bc={bc} S={s} C={c} H={h} Lt={LT} La={LA} f16", {}).get("dram_bytes")
    except (OSError, ValueError):
        pass
    return {"kernel": f"{name} via apx_ipattn_fwd ({launches_per_call} launch(es)/call), Bc={bc} S={s} C={c} H={h} Lt={LT} La={LA} fp16",
            "bound": bound, "achieved": round(achieved, 3), "peak": peak, "unit": unit, "frac": round(achieved / peak, 4),
            "traffic": traffic, "traffic_source": "profiles/ncu_traffic.json (ncu --set full capture of this launch; not re-measured by this run)",
            "us_per_call": round(us, 2), "algorithmic_bytes": byts, "algorithmic_flops": flops,
            "floor_us": round(1e6 * max(t_tensor, t_hbm), 3), "peak_source": pk["source"], "operand_copies_rotated": r}


def audiomae_timing(iters: int = 20) -> dict:
    """The other half of the hot path: AudioMAE ViT-B encode + pooling (apx_audiomae_fwd) for one 10 s clip and for a batch of 8,
    random-init weights, fp16 GEMM operands, CUDA-event timed over CUDA-graph replays.  Runs once per generation (2 forwards
    in the reference, pipeline_audioldm2.py:928-929), so it is reported beside the per-step metric, not inside it."""
    from ap_adapter_b200 import _lib
    from ap_adapter_b200.audiomae import AudioMAEConditionCTPoolRand

    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(0)
    enc = AudioMAEConditionCTPoolRand(checkpoint_path=None).to(dev).half().eval()
    out = {"kernels": "patchify + gemm16 (patch-embed: bias + pos / CLS residual) + 12 x (gemm16 qkv with norm1 folded, flash_attn (tcgen05, "
                      "N=513 keys in 128-key blocks, straight from the qkv slices), gemm_tc proj+res, gemm16 fc1+GELU with norm2 folded, "
                      "gemm_tc fc2+res) + layernorm + pool",
           "flops_per_clip": 97.0e9}
    for b in (1, 8):
        mel = torch.randn(b, 1024, 128, device=dev) * 0.45 - 0.45
        with torch.no_grad():
            for _ in range(2):
                enc(mel, time_pool=2, freq_pool=2)
            torch.cuda.synchronize()
            n0 = _lib.launch_count()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                emb = enc(mel, time_pool=2, freq_pool=2)[0]
            launches = _lib.launch_count() - n0
            graph.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
        assert torch.isfinite(emb).all()
        ms = e0.elapsed_time(e1) / iters
        # through the public module call: pinned-host mel in, host embedding out (the module replays its own CUDA graph)
        h_mel = mel.cpu().pin_memory()
        h_emb = torch.empty(emb.shape, dtype=emb.dtype).pin_memory()
        with torch.no_grad():
            for _ in range(3):
                h_emb.copy_(enc(h_mel.to(dev, non_blocking=True), time_pool=2, freq_pool=2)[0], non_blocking=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(iters):
                h_emb.copy_(enc(h_mel.to(dev, non_blocking=True), time_pool=2, freq_pool=2)[0], non_blocking=True)
                torch.cuda.current_stream().synchronize()
            e2e_ms = 1e3 * (time.perf_counter() - t0) / iters
        out[f"b{b}"] = {"ms": round(ms, 4), "clips_per_s": round(b / ms * 1e3, 1), "tflops": round(97.0e9 * b / (ms * 1e-3) / 1e12, 2),
                        "launches": launches, "e2e_ms_host_to_host": round(e2e_ms, 4)}
    return out


def clip_inputs(clips, dev):
    """Synthetic inputs of the given GLOBAL clip indices (one generator per clip, so any rank can rebuild any clip's inputs):
    latents (n, 8, 250, 16), [uncond; cond] 768-d context (2n, Lt + La, 768), T5 context (2n, 16, 1024), T5 mask (2n, 16)."""
    lat, g_u, g_c, p_u, p_c = [], [], [], [], []
    for c in clips:
        g = torch.Generator(device=dev).manual_seed(1000 + c)
        lat.append(torch.randn(1, *LAT_SHAPE, generator=g, device=dev, dtype=torch.float16))
        ctx = torch.randn(2, LT + LA, 768, generator=g, device=dev, dtype=torch.float16)
        t5 = torch.randn(2, L_T5, 1024, generator=g, device=dev, dtype=torch.float16)
        g_u.append(ctx[:1]); g_c.append(ctx[1:]); p_u.append(t5[:1]); p_c.append(t5[1:])
    n = len(lat)
    return (torch.cat(lat), torch.cat(g_u + g_c).contiguous(), torch.cat(p_u + p_c).contiguous(), torch.ones(2 * n, L_T5, device=dev))


def shard_equality_check(den, graph, gathered, b, world, rank, dev, ts, coefs, steps) -> dict:
    """SURVEY.md §4 item 4 on hardware: the gathered latents of the N-rank run must equal what ONE rank computes for the same clips.
    Rank 0 rebuilds the inputs of the LAST rank's clips, runs the same `steps` denoise steps through its own step graph and compares
    them with the rows the all-gather delivered (every rank runs the comparison on the shard after its own, so a bad rank cannot hide)."""
    import torch.distributed as dist
    other = (rank + 1) % world
    lat0, gpe, pe, mask = clip_inputs(range(other * b, (other + 1) * b), dev)
    g2 = den._step_graph(lat0, gpe, pe, mask, GUIDANCE)
    g2["latents"].copy_(lat0)
    for i in range(steps):
        g2["t"].copy_(ts[i % INFER_STEPS]); g2["coef"].copy_(coefs[i % INFER_STEPS])
        g2["graph"].replay()
        g2["latents"].copy_(g2["out"])
    mine = g2["latents"].float()
    theirs = gathered[other * b:(other + 1) * b].float()
    diff = ((mine - theirs).abs().max() / theirs.abs().max().clamp_min(1e-30)).item()
    stat = torch.tensor([diff, 0.0 if torch.equal(mine, theirs) else 1.0], device=dev, dtype=torch.float64)
    dist.all_reduce(stat, op=dist.ReduceOp.MAX)
    ok = stat[0].item() <= 1e-3
    assert ok, f"sharded run differs from the single-rank recomputation: max rel diff {stat[0].item():.3e}"
    return {"what": f"every rank recomputed the next rank's {b} clip(s) for {steps} steps and compared with the all-gathered rows",
            "max_rel_diff": stat[0].item(), "bitwise_equal_on_all_ranks": stat[1].item() == 0.0, "bar": 1e-3}


def run_config3(b, unet, dev, iters=1) -> dict:
    """BASELINE configs[2]: mel (host) -> AudioMAE(mel) + AudioMAE(0) -> context -> K/V caches -> 200 graphed CFG/DDIM steps ->
    latents (host), batch `b`, through APAdapterDenoiser.__call__.  AudioMAE is random-init (no checkpoint offline)."""
    from ap_adapter_b200.audiomae import AudioMAEConditionCTPoolRand
    from ap_adapter_b200.pipeline import APAdapterDenoiser
    torch.manual_seed(0)
    mae = AudioMAEConditionCTPoolRand(checkpoint_path=None).to(dev).eval()
    den = APAdapterDenoiser(unet, mae)
    g = torch.Generator().manual_seed(5)
    mel = (torch.randn(1, 1024, 128, generator=g) * 0.45 - 0.45).pin_memory()
    gpe8 = torch.randn(2 * b, LT, 768, generator=g).half().pin_memory()
    pe = torch.randn(2 * b, L_T5, 1024, generator=g).half().pin_memory()
    mask = torch.ones(2 * b, L_T5).pin_memory()
    lat = torch.randn(b, *LAT_SHAPE, generator=g).half().pin_memory()
    out_h = torch.empty(b, *LAT_SHAPE, dtype=torch.float16).pin_memory()

    def generate():
        o = den(mel.to(dev, non_blocking=True), pe.to(dev, non_blocking=True), mask.to(dev, non_blocking=True),
                gpe8.to(dev, non_blocking=True), time_pooling=2, freq_pooling=2, num_inference_steps=INFER_STEPS, guidance_scale=GUIDANCE,
                audio_length_in_s=10.0, latents=lat.to(dev, non_blocking=True))
        out_h.copy_(o, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    with torch.no_grad():
        t0 = time.perf_counter()
        generate()                                   # first call: captures the step graph, builds caches
        first = time.perf_counter() - t0
        times = []
        for _ in range(iters):
            den.invalidate_graph()                   # a NEW prompt: context caches + graph are rebuilt, as for every new generation
            t0 = time.perf_counter()
            generate()
            times.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        generate()                                   # the same prompt again: graph and caches are reused
        reuse = time.perf_counter() - t0
    assert torch.isfinite(out_h).all()
    sec = sum(times) / len(times)
    return {"clips": b, "ddim_steps": INFER_STEPS, "time_pooling": 2, "freq_pooling": 2, "seconds_per_generation": round(sec, 4),
            "steps_per_s": round(b * INFER_STEPS / sec, 2), "seconds_first_call": round(first, 3),
            "seconds_per_generation_graph_reused": round(reuse, 4), "steps_per_s_graph_reused": round(b * INFER_STEPS / reuse, 2),
            "what": "wall clock, host mel/embeddings/latents in, host latents out; includes 2 AudioMAE encodes on the first use of a pooling, "
                    "32 context K/V cache builds, graph capture (new prompt) and 200 replays"}


def run_train(args) -> dict:
    """BASELINE configs[4]: one optimiser step of adapter training per 'step' (see training.train_step), data parallel."""
    import torch.distributed as dist

    from ap_adapter_b200 import _lib, training
    from ap_adapter_b200.attention_processor import IPAttnProcessor2_0
    from ap_adapter_b200.pipeline import DDIMScheduler
    from ap_adapter_b200.unet import AudioLDM2UNet2DConditionModel, build_attn_processors

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    b = args.batch
    torch.manual_seed(0)
    unet = AudioLDM2UNet2DConditionModel().eval().to(dev, torch.bfloat16)           # frozen UNet in the mixed-precision dtype
    procs = build_attn_processors(unet, IPAttnProcessor2_0, ap_scale=1.0, device=dev, dtype=torch.float32)   # fp32 adapter weights
    unet.set_attn_processor(dict(procs))
    params = training.freeze_all_but_adapter(unet)
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-2)
    red = training.BucketedGradReducer(params) if world > 1 else None
    sched = DDIMScheduler()
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    lat = torch.randn(b, *LAT_SHAPE, generator=g, device=dev, dtype=torch.bfloat16)
    la_cycle = [512, 128, 32, 8]                                                   # pooling rate drawn from [1, 2, 4, 8] (train_apadapter_v2.py:445-446)
    ctxs = [torch.randn(b, LT + la, 768, generator=g, device=dev, dtype=torch.bfloat16) for la in la_cycle]
    pe = torch.randn(b, L_T5, 1024, generator=g, device=dev, dtype=torch.bfloat16)
    mask = torch.ones(b, L_T5, device=dev)

    def step(i):
        return training.train_step(unet, sched, opt, lat, ctxs[i % 4], pe, mask, generator=g, params=params, reducer=red)

    for i in range(max(3, args.warmup)):
        step(i)
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    loss = 0.0
    for i in range(args.steps):
        loss = step(i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        w = torch.cat([p.detach().reshape(-1)[:1024] for p in params])
        ws = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(ws, w)
        assert all(torch.equal(ws[0], x) for x in ws), "replicas diverged"
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return None
    return {"metric": "adapter_train_steps_per_s", "value": args.steps / (ms * 1e-3), "unit": "optimiser steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic (random-init UNet / adapter weights, N(0,1) latents and contexts)",
            "config": dict(workload_config(b, world), effective_batch=b * world, audio_tokens="cycled 512/128/32/8",
                           all_reduce_bytes_per_step=sum(p.numel() for p in params) * 4 if world > 1 else 0,
                           all_reduce="bucketed (25 MB), issued from autograd hooks, overlapped with the backward (training.BucketedGradReducer)"),
            "impl": "ours", "gpu_launches": launches, "final_loss": loss, "replicas_bitwise_equal": True if world > 1 else None}


def run_ours(args) -> dict:
    import torch.distributed as dist

    from ap_adapter_b200 import _lib
    from ap_adapter_b200.attention_processor import IPAttnProcessor2_0
    from ap_adapter_b200.pipeline import APAdapterDenoiser, gather_latents
    from ap_adapter_b200.unet import AudioLDM2UNet2DConditionModel, build_attn_processors

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: ap_adapter_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    assert lib.apx_device_ok() == 1, lib.apx_last_error().decode()

    b = args.batch
    torch.manual_seed(0)
    unet = AudioLDM2UNet2DConditionModel().eval().half().to(dev)
    ours = build_attn_processors(unet, IPAttnProcessor2_0, ap_scale=AP_SCALE, device=dev, dtype=torch.float16)
    unet.set_attn_processor(dict(ours))
    den = APAdapterDenoiser(unet)
    lat0, gpe, pe, mask = clip_inputs(range(rank * b, (rank + 1) * b), dev)
    den.scheduler.set_timesteps(INFER_STEPS, device=dev)
    ts, coefs = den.scheduler.timesteps, den.scheduler.coefficients(dev)

    with torch.no_grad():
        graph = den._step_graph(lat0, gpe, pe, mask, GUIDANCE)  # 2 eager warm-up steps + 1 captured step
        per_step_launches = graph["apx_launches_per_step"]

        def step(i):
            graph["t"].copy_(ts[i % INFER_STEPS])
            graph["coef"].copy_(coefs[i % INFER_STEPS])
            graph["graph"].replay()
            graph["latents"].copy_(graph["out"])

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---- warm-up + clocks ---------------------------------------------------------------------------------------
        # Clocks under load: nvidia-smi polling costs the polled GPU several per cent (round 1: the sampled loop read 8 % slower than
        # the unsampled one), so the sampler runs over an IDENTICAL UNTIMED repeat of the loop, not over the timed regions.  The repeat
        # runs BEFORE them: a timed loop that starts 50 ms after the first launch reads 3 - 6 % slower than the same loop a second
        # later (profiles/r2k: 9.74 ms/step over 50 steps, 9.41 over 400), so the >= 1.5 s repeat doubles as the warm-up.
        graph["latents"].copy_(lat0)
        t_load0 = time.perf_counter()
        for i in range(max(3, args.warmup)):
            step(i)
        barrier()
        w_est0 = time.perf_counter()
        for i in range(20):
            step(i)
        barrier()
        ms_est = (time.perf_counter() - w_est0) * 1e3 / 20
        graph["latents"].copy_(lat0)
        with ClockSampler(local) as clk:
            barrier()
            w0 = time.perf_counter()
            for i in range(max(args.steps, int(1500 / max(ms_est, 1e-3)))):   # >= 1.5 s under load: several 200 ms samples
                step(i)
            barrier()
            w1 = time.perf_counter()
        clocks = clk.summary(w0, w1)
        clocks["sampled"] = "during an identical untimed repeat of the timed loop, run right before it (polling perturbs the polled GPU)"
        # Settle: on this pool's boxes the first ~10 s of sustained load run ~6 % slow whatever the loop is (replay only, replay +
        # device copies, end to end: profiles/r2y_step_loops.txt — 9.28 ms/step for the first ~1000 steps, then 8.70 from one
        # 200-step loop to the next, nvidia-smi reporting 1965 MHz throughout), so a timed loop placed 3 s after start-up reads
        # slower than the end-to-end loop behind it.  Untimed 100-step blocks run until the step change has been seen (a block
        # >= 3 % faster than the first, confirmed by the next) or the device has been under load for 14 s; never past 25 s.
        blocks = []
        sb0, sb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        while True:
            sb0.record()
            for i in range(100):
                step(i)
            sb1.record()
            barrier()
            flags = torch.tensor([sb0.elapsed_time(sb1), time.perf_counter() - t_load0], device=dev, dtype=torch.float64)
            if world > 1:   # every rank takes the same decision (the barrier above is collective)
                dist.all_reduce(flags, op=dist.ReduceOp.MAX)
            blocks.append(float(flags[0].item()))
            under_load = float(flags[1].item())
            agree = len(blocks) >= 2 and abs(blocks[-1] - blocks[-2]) <= 0.01 * blocks[-2]
            if (agree and (blocks[-1] <= 0.97 * blocks[0] or under_load >= 14.0)) or under_load >= 25.0:
                break
        clocks["settle"] = {"steps": 100 * len(blocks), "first_block_ms_per_step": blocks[0] / 100, "last_block_ms_per_step": blocks[-1] / 100,
                            "seconds_under_load": round(under_load, 1)}

        # ---- device-resident timing -----------------------------------------------------------------
        graph["latents"].copy_(lat0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(args.steps):
            step(i)
        final = gather_latents(graph["latents"], world * b) if world > 1 else graph["latents"]
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        assert torch.isfinite(final).all(), "non-finite latents"
        shard_check = shard_equality_check(den, graph, final, b, world, rank, dev, ts, coefs, args.steps) if world > 1 else None

        # ---- end to end: the step's latents come from and go back to pinned host memory every step ------------------
        h_lat = lat0.cpu().pin_memory()
        h_t = ts.cpu().pin_memory()
        h_coef = coefs.cpu().pin_memory()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def e2e_step(i):
            graph["latents"].copy_(h_lat, non_blocking=True)
            graph["t"].copy_(h_t[i % INFER_STEPS], non_blocking=True)
            graph["coef"].copy_(h_coef[i % INFER_STEPS], non_blocking=True)
            graph["graph"].replay()
            h_lat.copy_(graph["out"], non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the caller reads x_{t-1} before issuing the next step

        for i in range(3):
            e2e_step(i)
        h_lat.copy_(lat0)
        barrier()
        e2.record()
        for i in range(args.steps):
            e2e_step(i)
        e3.record()
        barrier()
        ms_e2e = e2.elapsed_time(e3)
        assert torch.isfinite(h_lat).all()

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    lat_bytes = h_lat.numel() * h_lat.element_size()
    out = {
        "metric": METRIC, "value": world * b * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic (random-init UNet / adapter weights, N(0,1) latents and contexts)",
        "config": workload_config(b, world), "impl": "ours", "host_graph": HOST_GRAPH,
        "e2e": {"value": world * b * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": lat_bytes + 8 + 16,
                "d2h_bytes_per_step": lat_bytes, "ms_per_step": ms_e2e / args.steps,
                "path": "APAdapterDenoiser step graph: pinned-host latents/timestep/DDIM coefficients -> HBM, replay, x_{t-1} -> pinned host, sync"},
        "gpu_launches": per_step_launches * args.steps, "gpu_launches_per_step": per_step_launches,
        "clocks": clocks,
    }
    if shard_check is not None:
        out["shard_check"] = shard_check
    if rank == 0 and WORKLOAD == "step":
        out["roofline"] = site_roofline(b)
        try:
            out["roofline_gemm"] = gemm_roofline(b)
        except Exception as exc:  # noqa: BLE001
            out["roofline_gemm"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        out["audiomae"] = audiomae_timing()
        try:   # the fair same-box comparison; informational, never allowed to break the bench line
            out["torch_gpu"] = torch_gpu_step(unet, ours, lat0, gpe, pe, mask, ts, coefs, steps=min(args.steps, 50))
        except Exception as exc:  # noqa: BLE001
            out["torch_gpu"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        if world == 1 and not args.no_cpu_baseline and not args.no_extras:
            try:   # BASELINE configs[2] as one path (informational beside the headline; its own line: --workload config3)
                out["config3"] = run_config3(4, unet, dev, iters=1)
            except Exception as exc:  # noqa: BLE001
                out["config3"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_reference(b, steps=3, warmup=1, budget_s=25.0)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out if rank == 0 else None



class _TorchIPAttnSite(torch.nn.Module):
    """Bench-local PyTorch restatement of the reference adapter processor (attention_processor.py:347-470: to_q, SDPA over the
    text tokens, SDPA over the audio tokens through to_k_ip / to_v_ip, h + scale * h_ip, to_out) used ONLY for the `torch_gpu`
    comparison leg: the same UNet step with stock cuBLAS + SDPA kernels on the same GPU.  Never on the product path."""

    def __init__(self, hidden_size, name=None, cross_attention_dim=None, num_tokens=4, scale=1.0, do_copy=False):
        super().__init__()
        self.num_tokens, self.scale = num_tokens, scale
        self.to_k_ip = torch.nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)
        self.to_v_ip = torch.nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0):
        import torch.nn.functional as F
        b, _, c = hidden_states.shape
        h = attn.heads

        def split(t):
            return t.view(b, -1, h, c // h).transpose(1, 2)

        q = split(attn.to_q(hidden_states))
        e_t, e_a = encoder_hidden_states[:, :self.num_tokens], encoder_hidden_states[:, self.num_tokens:]
        o = F.scaled_dot_product_attention(q, split(attn.to_k(e_t)), split(attn.to_v(e_t)))
        o_ip = F.scaled_dot_product_attention(q, split(self.to_k_ip(e_a)), split(self.to_v_ip(e_a)))
        o = (o + self.scale * o_ip).transpose(1, 2).reshape(b, -1, c)
        return attn.to_out[1](attn.to_out[0](o))


def torch_gpu_step(unet, ours, lat0, gpe, pe, mask, ts, coefs, steps: int) -> dict:
    """The same denoise step with the stock PyTorch graph on the same GPU: reference-algorithm processors (cuBLAS + SDPA, fp16),
    unfused FeedForward / projections / residual adds, captured into a CUDA graph and timed exactly like the product step."""
    from ap_adapter_b200.attention_processor import AttnProcessor2_0
    from ap_adapter_b200.pipeline import APAdapterDenoiser
    from ap_adapter_b200.unet import NHWC, BasicTransformerBlock, FeedForward, build_attn_processors
    saved = (FeedForward.use_fused, AttnProcessor2_0.fuse_projections, BasicTransformerBlock.fold_residual, NHWC.enabled, AttnProcessor2_0.cache_context,
             AttnProcessor2_0.flash_min_keys, BasicTransformerBlock.add_ln)
    dev = lat0.device
    try:
        FeedForward.use_fused = AttnProcessor2_0.fuse_projections = BasicTransformerBlock.fold_residual = NHWC.enabled = False
        AttnProcessor2_0.cache_context = False
        AttnProcessor2_0.flash_min_keys = 1 << 30     # the stock graph keeps torch's SDPA everywhere
        BasicTransformerBlock.add_ln = False          # ... and ATen's residual adds / LayerNorms
        theirs = build_attn_processors(unet, _TorchIPAttnSite, ap_scale=AP_SCALE, device=dev, dtype=torch.float16)
        for k, p in ours.items():
            if hasattr(p, "to_k_ip"):
                theirs[k].load_state_dict(p.state_dict())
        unet.set_attn_processor(dict(theirs))
        den = APAdapterDenoiser(unet)
        with torch.no_grad():
            graph = den._step_graph(lat0, gpe, pe, mask, GUIDANCE)
            assert graph["apx_launches_per_step"] == 0

            def step(i):
                graph["t"].copy_(ts[i % INFER_STEPS])
                graph["coef"].copy_(coefs[i % INFER_STEPS])
                graph["graph"].replay()
                graph["latents"].copy_(graph["out"])

            graph["latents"].copy_(lat0)
            for i in range(3):
                step(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for i in range(steps):
                step(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
        return {"value": lat0.shape[0] * 1e3 / ms, "unit": UNIT, "ms_per_step": ms, "steps_timed": steps,
                "what": "same step, stock PyTorch graph on the same GPU: reference-algorithm processors (cuBLAS + cuDNN SDPA, fp16), "
                        "unfused FeedForward / projections / residual adds, one CUDA graph per step, device-timed"}
    finally:
        (FeedForward.use_fused, AttnProcessor2_0.fuse_projections, BasicTransformerBlock.fold_residual, NHWC.enabled,
         AttnProcessor2_0.cache_context, AttnProcessor2_0.flash_min_keys, BasicTransformerBlock.add_ln) = saved
        unet.set_attn_processor(dict(ours))


def run_config3_line(args) -> dict:
    from ap_adapter_b200.attention_processor import IPAttnProcessor2_0
    from ap_adapter_b200.unet import AudioLDM2UNet2DConditionModel, build_attn_processors
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    unet = AudioLDM2UNet2DConditionModel().eval().half().to(dev)
    unet.set_attn_processor(dict(build_attn_processors(unet, IPAttnProcessor2_0, ap_scale=AP_SCALE, device=dev, dtype=torch.float16)))
    r = run_config3(args.batch, unet, dev, iters=2)
    return {"metric": METRIC, "value": r["steps_per_s"], "unit": UNIT, "n_gpus": 1, "steps": INFER_STEPS, "warmup": 1,
            "ms_per_step": 1e3 * r["seconds_per_generation"] / INFER_STEPS, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic (random-init UNet / AudioMAE / adapter weights, N(-0.45, 0.45) mel, N(0,1) latents and contexts)",
            "config": workload_config(args.batch, 1), "impl": "ours", "e2e": {"value": r["steps_per_s"], "unit": UNIT,
            "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "note": "host tensors in / out once per GENERATION (200 steps), see config3"},
            "config3": r}


def run_reference(args) -> dict:
    """The reference's CPU path on the host cores.  Under torchrun only rank 0 works; it then runs the WHOLE job's clips (n_gpus x
    clips_per_gpu) so that its line describes the same workload as the N-GPU arm (the host does not get faster with more GPUs)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return None
    if args.workload not in ("step", "config4"):
        return {"impl": "reference", "metric": METRIC, "value": None, "unit": UNIT, "n_gpus": args.gpus,
                "declined": f"the CPU arm times the denoise step (workloads step / config4), not '{args.workload}'"}
    n = max(1, args.gpus)
    r = cpu_reference(args.batch * n, steps=args.steps, warmup=max(1, min(args.warmup, 2)), budget_s=150.0)
    return {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps_timed"],
            "warmup": 1, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (random-init UNet / adapter weights, N(0,1) latents and contexts)",
            "config": dict(workload_config(args.batch, n), host_arm=f"one host process runs all {args.batch * n} clip(s) of the {n}-GPU job"),
            "impl": "reference",
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "processors", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="clips per GPU (CFG doubles it inside the UNet)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational config3 leg of the default workload")
    ap.add_argument("--workload", default="step", choices=["step", "config3", "config4", "train"])
    args = ap.parse_args()
    set_workload(args.workload, args)
    if args.impl == "reference":
        out = run_reference(args)
    elif args.workload == "train":
        out = run_train(args)
    elif args.workload == "config3":
        out = run_config3_line(args)
    else:
        out = run_ours(args)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
