"""AudioMAE ViT-B encoder + time/frequency token pooling on the sm_100a kernels.

Mirrors the reference interface in ``audio_encoder/AudioMAE.py``:

* ``Vanilla_AudioMAE``            (:17-55)   encoder wrapper, ``forward(x, mask_ratio, no_mask, no_average)``
* ``AudioMAEConditionCTPoolRand`` (:94-212)  ``model(mel[B,1024,128], time_pool=, freq_pool=) -> [emb, ones]``
* ``extract_kaldi_fbank_feature`` (:356-394) waveform -> normalised (1024, 128) log-mel

The parameter tree uses the reference's state-dict names (``patch_embed.proj.weight``, ``cls_token``, ``pos_embed``,
``blocks.N.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}``, ``norm``; ``audio_encoder/models_mae.py:199-231``), so a
real ``pretrained.pth`` drops in.  The arithmetic (``models_mae.py:548-570`` + ``AudioMAE.py:148-182``) runs in
``apx_audiomae_fwd``: fp32 residual stream / LayerNorm / softmax statistics, fp16 (or bf16) tensor-core GEMM operands.
Only the encoder is built; the MAE decoder, masking and losses are unused at inference.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib

EMBED, DEPTH, HEADS, MLP, NTOK = 768, 12, 12, 3072, 513


def sincos_pos_embed(embed_dim: int = EMBED, grid_size=(8, 64)) -> torch.Tensor:
    """The frozen 2-D sin-cos table the reference builds at init (``models_mae.py:87-151,330-340``) on the
    (freq=8, time=64) ``patch_hw`` grid, with a zero row for CLS.  Real checkpoints carry it as ``pos_embed``."""
    def one_d(dim, pos):
        omega = 1.0 / 10000 ** (np.arange(dim // 2, dtype=float) / (dim / 2.0))
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    gw, gh = np.meshgrid(np.arange(grid_size[1], dtype=np.float32), np.arange(grid_size[0], dtype=np.float32))
    emb = np.concatenate([one_d(embed_dim // 2, gw), one_d(embed_dim // 2, gh)], axis=1)
    emb = np.concatenate([np.zeros([1, embed_dim]), emb], axis=0)
    return torch.from_numpy(emb).float().unsqueeze(0)


class _Attention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, hidden)


class _PatchEmbed(nn.Module):
    def __init__(self):
        super().__init__()
        self.proj = nn.Conv2d(1, EMBED, kernel_size=16, stride=16)


class MaskedAutoencoderViTEncoder(nn.Module):
    """Parameter container with the encoder half of ``mae_vit_base_patch16(in_chans=1, audio_exp=True,
    img_size=(1024,128))`` (``models_mae.py:689-701``); the forward is the fused kernel sequence."""

    # norm1 / norm2 of every block folded into the qkv / fc1 GEMMs (24 fewer kernels per encode); APX_MAE_LN_FOLD=0 keeps them separate
    fold_norms = __import__("os").environ.get("APX_MAE_LN_FOLD", "1") != "0"

    def __init__(self, depth: int = DEPTH):
        super().__init__()
        self.patch_embed = _PatchEmbed()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, EMBED))
        self.pos_embed = nn.Parameter(sincos_pos_embed(), requires_grad=False)
        self.blocks = nn.ModuleList([_Block(EMBED, MLP) for _ in range(depth)])
        self.norm = nn.LayerNorm(EMBED, eps=1e-6)
        w = self.patch_embed.proj.weight.data
        torch.nn.init.xavier_uniform_(w.view([w.shape[0], -1]))  # models_mae.py:355-357
        torch.nn.init.normal_(self.cls_token, std=0.02)
        self._pack = None

    # ---- device weight pack for the C ABI -------------------------------------------------------------
    def _compute_dtype(self) -> torch.dtype:
        dt = self.blocks[0].attn.qkv.weight.dtype
        return dt if dt in (torch.float16, torch.bfloat16) else torch.float16

    def _weight_pack(self):
        params = list(self.parameters())
        key = tuple((p.data_ptr(), p._version, p.dtype) for p in params)
        if self._pack is not None and self._pack[0] == key:
            return self._pack[1]
        dev = self.cls_token.device
        if dev.type != "cuda":
            raise _lib.ApxError("AudioMAE encoder weights are on the CPU; call .cuda() (no CPU fallback)")
        cdt = self._compute_dtype()
        keep = []  # device tensors the raw pointers refer to

        def f32(t):
            t = t.detach().float().contiguous()
            keep.append(t)
            return t.data_ptr()

        def w16(t):
            t = t.detach().to(cdt).contiguous()
            keep.append(t)
            return t.data_ptr()

        def folded(lin, norm):
            # norm folded into lin (apx_linear_lnfold): Wg = W * gamma in the compute dtype, c1 = row sums of Wg, c2 = W beta + b
            wg, c1, c2 = _lib.fold_layernorm(lin.weight.detach().to(cdt), lin.bias.detach(), norm.weight.detach(), norm.bias.detach())
            keep.extend((wg, c1, c2))
            return wg.data_ptr(), c1.data_ptr(), c2.data_ptr()

        blocks = (_lib.VitBlock * len(self.blocks))()
        for i, b in enumerate(self.blocks):
            fold = (folded(b.attn.qkv, b.norm1) + folded(b.mlp.fc1, b.norm2)) if self.fold_norms else (None,) * 6
            blocks[i] = _lib.VitBlock(
                f32(b.norm1.weight), f32(b.norm1.bias), w16(b.attn.qkv.weight), f32(b.attn.qkv.bias),
                w16(b.attn.proj.weight), f32(b.attn.proj.bias), f32(b.norm2.weight), f32(b.norm2.bias),
                w16(b.mlp.fc1.weight), f32(b.mlp.fc1.bias), w16(b.mlp.fc2.weight), f32(b.mlp.fc2.bias), *fold)
        patch_wT = self.patch_embed.proj.weight.detach().float().reshape(EMBED, 256).t().contiguous()
        keep.append(patch_wT)
        w = _lib.VitWeights(len(self.blocks), EMBED, HEADS, MLP, _lib.dtype_code(cdt), patch_wT.data_ptr(),
                            f32(self.patch_embed.proj.bias), f32(self.pos_embed.reshape(NTOK, EMBED)),
                            f32(self.cls_token.reshape(EMBED)), f32(self.norm.weight), f32(self.norm.bias), blocks)
        self._pack = (key, (w, blocks, keep))
        return self._pack[1]

    # the ~80 launches of one encode are replayed from a CUDA graph (per batch size / pooling / weight version); APX_MAE_GRAPH=0 disables
    use_graph = __import__("os").environ.get("APX_MAE_GRAPH", "1") != "0"

    def _encode_pooled_eager(self, mel, emb, time_pool, freq_pool):
        lib = _lib.load()
        w, _, _ = self._weight_pack()
        dev = mel.device
        b = mel.shape[0]
        wsb = lib.apx_audiomae_workspace_bytes(b)
        work = _lib.workspace(wsb, dev, "audiomae")
        with torch.cuda.device(dev):
            _lib.check(lib.apx_audiomae_fwd(_lib.ptr(mel), ctypes.byref(w), _lib.ptr(emb), b, int(time_pool), int(freq_pool),
                                            _lib.ptr(work), wsb, _lib.stream_ptr()), "apx_audiomae_fwd")

    def encode_pooled(self, mel: torch.Tensor, time_pool: int, freq_pool: int) -> torch.Tensor:
        """mel (B, 1024, 128) fp32 on the module's device -> (B, 512/(tp*fp), 768) fp32."""
        dev = self.cls_token.device
        if dev.type != "cuda":
            raise _lib.ApxError("AudioMAE encoder weights are on the CPU; call .cuda() (no CPU fallback)")
        mel = mel.to(device=dev, dtype=torch.float32).contiguous()
        b = mel.shape[0]
        ntok = (64 // time_pool) * (8 // freq_pool)
        if not self.use_graph or torch.cuda.is_current_stream_capturing():
            emb = torch.empty((b, ntok, EMBED), dtype=torch.float32, device=dev)
            self._encode_pooled_eager(mel, emb, time_pool, freq_pool)
            return emb
        self._weight_pack()
        key = (b, int(time_pool), int(freq_pool), self._pack[0])
        g = self.__dict__.get("_graph")
        if g is None or g["key"] != key:
            g = dict(key=key, mel=torch.empty_like(mel), emb=torch.empty((b, ntok, EMBED), dtype=torch.float32, device=dev), ws={})
            g["mel"].copy_(mel)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                      # warm-up off the capture (also validates the arguments)
                self._encode_pooled_eager(g["mel"], g["emb"], time_pool, freq_pool)
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with _lib.workspace_scope(g["ws"]), torch.cuda.graph(graph):
                self._encode_pooled_eager(g["mel"], g["emb"], time_pool, freq_pool)
            g["graph"] = graph
            self.__dict__["_graph"] = g
        g["mel"].copy_(mel)
        g["graph"].replay()
        return g["emb"].clone()

    def forward_encoder_no_random_mask_no_average(self, x: torch.Tensor) -> torch.Tensor:
        """(B,1,1024,128) -> (B,513,768), the un-pooled encoder output (``models_mae.py:548-570``)."""
        lib = _lib.load()
        w, _, _ = self._weight_pack()
        dev = self.cls_token.device
        mel = x.reshape(-1, 1024, 128).to(device=dev, dtype=torch.float32).contiguous()
        b = mel.shape[0]
        tokens = torch.empty((b, NTOK, EMBED), dtype=torch.float32, device=dev)
        wsb = lib.apx_audiomae_workspace_bytes(b)
        work = _lib.workspace(wsb, dev, "audiomae")
        with torch.cuda.device(dev):
            _lib.check(lib.apx_audiomae_encode(_lib.ptr(mel), ctypes.byref(w), _lib.ptr(tokens), b, _lib.ptr(work), wsb,
                                               _lib.stream_ptr()), "apx_audiomae_encode")
        return tokens


class Vanilla_AudioMAE(nn.Module):
    """Audio Masked Autoencoder (MAE) pre-trained on AudioSet (for AudioLDM2) — reference ``AudioMAE.py:17-55``.

    ``Vanilla_AudioMAE()`` loads ``pretrained.pth`` from the working directory exactly as the reference does;
    ``state_dict=`` / ``checkpoint_path=`` are additions for tests and benchmarks (no checkpoint exists offline).
    """

    def __init__(self, checkpoint_path: str | None = "pretrained.pth", state_dict=None, depth: int = DEPTH):
        super().__init__()
        model = MaskedAutoencoderViTEncoder(depth=depth)
        if state_dict is None and checkpoint_path is not None:
            from .weights import load_audiomae_checkpoint
            state_dict = load_audiomae_checkpoint(checkpoint_path)
        if state_dict is not None:
            enc = {k: v for k, v in state_dict.items() if not k.startswith(("decoder", "mask_token"))}
            model.load_state_dict(enc, strict=False)
        self.model = model.eval()

    def forward(self, x, mask_ratio=0.0, no_mask=False, no_average=False):
        """x: mel fbank [Batch, 1, 1024 (T), 128 (F)] -> embed [B, 513, 768] (no_mask=True, no_average=True)."""
        with torch.no_grad():
            if no_mask and no_average:
                return self.model.forward_encoder_no_random_mask_no_average(x)
            raise RuntimeError("This function is deprecated")  # as in the reference for every other combination


class AudioMAEConditionCTPoolRand(nn.Module):
    """Reference ``AudioMAE.py:94-212``: unmasked AudioMAE encode, drop CLS, (avg+max)/2 pool of the 64x8 token grid."""

    def __init__(self, time_pooling_factors=[1, 2, 4, 8], freq_pooling_factors=[1, 2, 4, 8], eval_time_pooling=8,
                 eval_freq_pooling=8, mask_ratio=0.0, regularization=False, no_audiomae_mask=True,
                 no_audiomae_average=True, checkpoint_path: str | None = "pretrained.pth", state_dict=None,
                 depth: int = DEPTH):
        super().__init__()
        self.device = None
        self.time_pooling_factors = time_pooling_factors
        self.freq_pooling_factors = freq_pooling_factors
        self.no_audiomae_mask = no_audiomae_mask
        self.no_audiomae_average = no_audiomae_average
        self.eval_freq_pooling = eval_freq_pooling
        self.eval_time_pooling = eval_time_pooling
        self.mask_ratio = mask_ratio
        self.use_reg = regularization
        self.audiomae = Vanilla_AudioMAE(checkpoint_path=checkpoint_path, state_dict=state_dict, depth=depth)
        self.audiomae.eval()
        for p in self.audiomae.parameters():
            p.requires_grad = False

    def get_unconditional_condition(self, batchsize):
        param = next(self.audiomae.parameters())
        device = param.device
        time_pool, freq_pool = min(self.eval_time_pooling, 64), min(self.eval_freq_pooling, 8)
        token_num = int(512 / (time_pool * freq_pool))
        return [torch.zeros((batchsize, token_num, 768)).to(device).float(),
                torch.ones((batchsize, token_num)).to(device).float()]

    def pool(self, representation, time_pool=None, freq_pool=None):
        """Pooling of an already-encoded (B,513,768) representation (``AudioMAE.py:148-182``) with the pool kernel."""
        assert representation.size(-1) == 768
        lib = _lib.load()
        rep = representation.to(torch.float32).contiguous()
        _lib.require_cuda(rep)
        b = rep.shape[0]
        ntok = (64 // time_pool) * (8 // freq_pool)
        emb = torch.empty((b, ntok, 768), dtype=torch.float32, device=rep.device)
        with torch.cuda.device(rep.device):
            _lib.check(lib.apx_pool_tokens(_lib.ptr(rep), _lib.ptr(emb), b, int(time_pool), int(freq_pool), _lib.stream_ptr()),
                       "apx_pool_tokens")
        return emb

    def regularization(self, x):
        assert x.size(-1) == 768
        return torch.nn.functional.normalize(x, p=2, dim=-1)

    def forward(self, batch, time_pool=None, freq_pool=None):
        assert batch.size(-2) == 1024 and batch.size(-1) == 128
        if self.device is None:
            self.device = next(self.audiomae.parameters()).device
        if not (self.no_audiomae_mask and self.no_audiomae_average):
            raise RuntimeError("This function is deprecated")
        with torch.no_grad():
            representation = self.audiomae.model.encode_pooled(batch.reshape(-1, 1024, 128), time_pool, freq_pool)
            pdt = next(self.audiomae.parameters()).dtype
            if pdt in (torch.float16, torch.bfloat16):
                representation = representation.to(pdt)  # a model cast with .to(dtype) returns that dtype
            if self.use_reg:
                representation = self.regularization(representation)
            # AudioMAE.py:207-212 builds the ones on the host and copies them over; created on the device here (same values,
            # no host->device copy, CUDA-graph capturable)
            return [representation,
                    torch.ones((representation.size(0), representation.size(1)), device=representation.device, dtype=torch.float32)]


def kaldi_mel_banks(num_bins: int = 128, padded_window: int = 512, sample_freq: float = 16000.0, low_freq: float = 20.0,
                    high_freq: float = 0.0) -> torch.Tensor:
    """The mel filterbank of Kaldi's fbank as torchaudio builds it (``compliance/kaldi.py::get_mel_banks``, no VTLN):
    triangular filters on the mel scale ``1127 ln(1 + f / 700)``; returns (num_bins, padded_window // 2 + 1) with a zero last
    column (the Nyquist bin is dropped)."""
    nyq = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyq
    mel = lambda f: 1127.0 * np.log(1.0 + np.asarray(f, dtype=np.float64) / 700.0)  # noqa: E731
    lo, hi = float(mel(low_freq)), float(mel(high_freq))
    delta = (hi - lo) / (num_bins + 1)
    b = np.arange(num_bins, dtype=np.float64)[:, None]
    left, center, right = lo + b * delta, lo + (b + 1.0) * delta, lo + (b + 2.0) * delta
    n_fft = padded_window // 2
    m = mel(sample_freq / padded_window * np.arange(n_fft, dtype=np.float64))[None, :]
    up, down = (m - left) / (center - left), (right - m) / (right - center)
    bins = np.maximum(0.0, np.minimum(up, down))
    return torch.from_numpy(np.pad(bins, ((0, 0), (0, 1))).astype(np.float32))


def sinc_resample_taps(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """The windowed-sinc polyphase table of ``torchaudio.functional.resample`` (default ``sinc_interp_hann``; torchaudio is a third-party
    dependency of the reference, AudioMAE.py:360-366): returns (taps fp32 [new, 2*width + orig], orig, new, width) with orig / new
    reduced by their gcd.  taps[p, t] = sinc-Hann weight of input sample ``t - width`` for output phase ``p``."""
    import math

    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, :] / orig
    t = torch.arange(0, -new, -1, dtype=torch.float64)[:, None] / new + idx
    t = (t * base).clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base / orig
    taps = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t) * window * scale
    return taps.float().contiguous(), orig, new, width


_MEL_FBT = {}
_TAPS = {}


def resample_to_16k(waveform: torch.Tensor, sampling_rate: int) -> torch.Tensor:
    """(B, n) or (n,) CUDA waveform -> 16 kHz with libapx's polyphase resampler (apx_resample); same taps as torchaudio."""
    if int(sampling_rate) == 16000:
        return waveform
    lib = _lib.load()
    dev = waveform.device
    w = waveform.reshape(-1, waveform.shape[-1]).float().contiguous()
    key = (dev, int(sampling_rate))
    if key not in _TAPS:
        taps, orig, new, width = sinc_resample_taps(int(sampling_rate), 16000)
        _TAPS[key] = (taps.to(dev), orig, new, width)
    taps, orig, new, width = _TAPS[key]
    b, n = w.shape
    n_out = -(-new * n // orig)
    out = torch.empty((b, n_out), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.apx_resample(_lib.ptr(w), b, n, _lib.ptr(taps), orig, new, width, _lib.ptr(out), _lib.stream_ptr()), "apx_resample")
    return out.reshape(waveform.shape[:-1] + (n_out,))


def extract_kaldi_fbank_feature(waveform, sampling_rate, log_mel_spec):
    """Reference ``AudioMAE.py:356-394``: resample to 16 kHz, remove DC, Kaldi fbank (128 mel bins, hanning, 10 ms
    shift), cut/pad to ``log_mel_spec.size(0)`` frames, normalise with the AudioSet statistics.

    A CUDA waveform runs through libapx: ``apx_resample`` (polyphase windowed-sinc, torchaudio's table) and ``apx_fbank`` (framing, DC
    removal, pre-emphasis, window, 512-point DFT, mel filterbank, log, normalisation); a CPU waveform takes the reference's own route
    (torchaudio on the host, which is where ``pipeline_audioldm2.py:919-925`` computes it)."""
    import torchaudio

    norm_mean = -4.2677393
    norm_std = 4.5689974
    target_len = log_mel_spec.size(0)
    if waveform.is_cuda:
        return extract_kaldi_fbank_feature_batch(waveform.reshape(-1, waveform.shape[-1])[:1], sampling_rate, target_len)[0]  # channel 0
    if sampling_rate != 16000:
        waveform_16k = torchaudio.functional.resample(waveform, orig_freq=sampling_rate, new_freq=16000)
    else:
        waveform_16k = waveform
    waveform_16k = waveform_16k - waveform_16k.mean()
    fbank = torchaudio.compliance.kaldi.fbank(waveform_16k, htk_compat=True, sample_frequency=16000, use_energy=False,
                                              window_type="hanning", num_mel_bins=128, dither=0.0, frame_shift=10)
    p = target_len - fbank.shape[0]
    if p > 0:
        fbank = torch.nn.functional.pad(fbank, (0, 0, 0, p))
    elif p < 0:
        fbank = fbank[:target_len, :]
    return (fbank - norm_mean) / (norm_std * 2)


def extract_kaldi_fbank_feature_batch(waveforms: torch.Tensor, sampling_rate: int, target_len: int = 1024) -> torch.Tensor:
    """The same front-end for a BATCH of equally long mono clips on the GPU (the training collate, train_apadapter_v2.py:426-468, calls
    the reference function clip by clip): (B, n) CUDA waveforms at ``sampling_rate`` -> (B, target_len, 128) normalised log-mel.
    Two launches for the resampling-free case (per-clip means + one block per (frame, clip)), three otherwise."""
    _lib.require_cuda(waveforms)
    lib = _lib.load()
    dev = waveforms.device
    wav = resample_to_16k(waveforms.reshape(-1, waveforms.shape[-1]).float().contiguous(), sampling_rate)
    b, n = wav.shape
    if dev not in _MEL_FBT:
        _MEL_FBT[dev] = kaldi_mel_banks().t().contiguous().to(dev)
    out = torch.empty((b, target_len, 128), dtype=torch.float32, device=dev)
    scratch = torch.empty(b, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.apx_fbank_batch(_lib.ptr(wav), b, n, _lib.ptr(_MEL_FBT[dev]), _lib.ptr(out), target_len, _lib.ptr(scratch),
                                       _lib.stream_ptr()), "apx_fbank_batch")
    return out


def read_wav(path: str):
    """(waveform (1, T) float32 in [-1, 1], sample_rate) via scipy — ``torchaudio.load`` needs torchcodec, which
    this image lacks."""
    from scipy.io import wavfile

    sr, wav = wavfile.read(path)
    if wav.dtype.kind == "i":
        wav = wav.astype(np.float32) / float(np.iinfo(wav.dtype).max + 1)
    wav = torch.from_numpy(np.asarray(wav, dtype=np.float32))
    wav = wav.unsqueeze(0) if wav.dim() == 1 else wav.t()[:1]
    return wav, int(sr)

