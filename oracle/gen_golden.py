"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (TEST INFRASTRUCTURE).

    python -m oracle.gen_golden            # needs /root/reference (read-only); writes tests/golden/

The reference has no tests or golden vectors of its own, so parity is pinned on outputs of the
reference's own modules executed here on the seeded cases of oracle/cases.py:

* ``IPAttnProcessor2_0`` / ``AttnProcessor2_0`` from ``APadapter/ap_adapter/attention_processor.py`` driven
  through a duck-typed ``StubAttention`` (oracle/ref_shims.py);
* ``mae_vit_base_patch16(in_chans=1, audio_exp=True, img_size=(1024,128))`` from ``audio_encoder/models_mae.py``
  -> ``forward_encoder_no_random_mask_no_average`` and ``AudioMAEConditionCTPoolRand.pool`` /
  ``extract_kaldi_fbank_feature`` from ``audio_encoder/AudioMAE.py``.

The fixtures travel to the GPU box; /root/reference does not.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import cases, ref_shims


def _sub(name: str, y: torch.Tensor) -> np.ndarray:
    y2 = y.reshape(-1, y.shape[-1])
    return y2[:: cases.ROW_STRIDE.get(name, 1)].contiguous().numpy()


def gen_real_weights():
    k = torch.load(os.path.join(ref_shims.REF_ROOT, "copied_cross_attention", cases.REAL_SITE + "_k.bin"), map_location="cpu", weights_only=False)
    v = torch.load(os.path.join(ref_shims.REF_ROOT, "copied_cross_attention", cases.REAL_SITE + "_v.bin"), map_location="cpu", weights_only=False)
    k, v = k.detach().to(torch.float16).numpy(), v.detach().to(torch.float16).numpy()
    assert k.shape == (256, 768) and v.shape == (256, 768)
    np.savez_compressed(os.path.join(cases.GOLDEN_DIR, "real_ip_weights.npz"), k=k, v=v, site=cases.REAL_SITE)


def gen_ipattn():
    ap = ref_shims.load_attention_processor()
    out = {}
    for name in cases.IPATTN_CASES:
        c = cases.make_ipattn_case(name)
        hidden = c["x"].shape[-1]
        attn = ref_shims.StubAttention(hidden, cases.DCTX, c["heads"])
        proc = ap.IPAttnProcessor2_0(hidden_size=hidden, name=name, cross_attention_dim=cases.DCTX,
                                     num_tokens=c["num_tokens"], scale=c["scale"], do_copy=False)
        with torch.no_grad():
            attn.to_q.weight.copy_(c["wq"]); attn.to_k.weight.copy_(c["wk"]); attn.to_v.weight.copy_(c["wv"])
            attn.to_out[0].weight.copy_(c["wo"]); attn.to_out[0].bias.copy_(c["bo"])
            proc.to_k_ip.weight.copy_(c["wk_ip"]); proc.to_v_ip.weight.copy_(c["wv_ip"])
            y = proc(attn, c["x"], encoder_hidden_states=c["ctx"])
        out[name] = _sub(name, y)
        print(f"ipattn {name}: out {tuple(y.shape)} -> stored {out[name].shape}, |y|max {y.abs().max():.3f}")
    for name in cases.ATTN_CASES:
        c = cases.make_attn_case(name)
        hidden = c["x"].shape[-1]
        attn = ref_shims.StubAttention(hidden, None if c["ctx"] is None else c["ctx"].shape[-1], c["heads"])
        proc = ap.AttnProcessor2_0()
        with torch.no_grad():
            attn.to_q.weight.copy_(c["wq"]); attn.to_k.weight.copy_(c["wk"]); attn.to_v.weight.copy_(c["wv"])
            attn.to_out[0].weight.copy_(c["wo"]); attn.to_out[0].bias.copy_(c["bo"])
            y = proc(attn, c["x"], encoder_hidden_states=c["ctx"], attention_mask=c["mask"])
        out["plain_" + name] = y.reshape(-1, y.shape[-1]).numpy()
        print(f"attn {name}: out {tuple(y.shape)}")
    np.savez_compressed(os.path.join(cases.GOLDEN_DIR, "ipattn_reference.npz"), **out)


def gen_audiomae():
    models_mae, audiomae = ref_shims.load_audio_encoder()
    torch.manual_seed(0)
    model = models_mae.mae_vit_base_patch16(in_chans=1, audio_exp=True, img_size=(1024, 128)).eval()
    out = {}
    # the reference's own frozen sin-cos table (models_mae.py:330-340), to pin the oracle's restatement of it
    out["pos_embed_rows"] = model.pos_embed[0, ::57].detach().numpy()
    sd = cases.mae_state_dict()
    assert torch.equal(sd["pos_embed"], model.pos_embed.detach()), "oracle sin-cos table differs from the reference's"
    missing = model.load_state_dict(sd, strict=False)
    assert not [k for k in missing.missing_keys if not k.startswith(("decoder", "mask_token"))], missing.missing_keys
    pool = audiomae.AudioMAEConditionCTPoolRand.pool

    class _Self:  # `pool` only assigns self.avgpooling / self.maxpooling
        pass

    # real input: the reference's own Kaldi fbank of its own piano.wav (torchaudio.load needs torchcodec -> scipy)
    from scipy.io import wavfile
    sr, wav = wavfile.read(os.path.join(ref_shims.REF_ROOT, "piano.wav"))
    wav = torch.from_numpy(wav.astype(np.float32) / 32768.0)
    wav = wav.unsqueeze(0) if wav.dim() == 1 else wav.t()[:1]
    fbank = audiomae.extract_kaldi_fbank_feature(wav, sr, torch.zeros((1024, 128)))
    mel_piano = fbank.to(torch.float16)
    out["mel_piano_f16"] = mel_piano.numpy()
    print(f"piano.wav: sr {sr}, fbank {tuple(fbank.shape)} mean {fbank.mean():.3f} std {fbank.std():.3f}")

    for tag, mel in (("synth", cases.make_mel()), ("piano", mel_piano.float().unsqueeze(0))):
        with torch.no_grad():
            rep = model.forward_encoder_no_random_mask_no_average(mel.unsqueeze(1))
            out[f"{tag}_encoder_rows"] = rep[0, ::8].numpy()          # 65 of the 513 rows
            for tp, fp in cases.POOLS:
                pooled = pool(_Self(), rep, tp, fp)
                out[f"{tag}_pool_{tp}x{fp}"] = (pooled[0, ::4] if pooled.shape[1] > 128 else pooled[0]).numpy()
        print(f"audiomae {tag}: encoder {tuple(rep.shape)} |rep|max {rep.abs().max():.3f}")
    np.savez_compressed(os.path.join(cases.GOLDEN_DIR, "audiomae_reference.npz"), **out)


def gen_frontend():
    """Waveform fixtures for the GPU front-end (apx_resample / apx_fbank), made with the reference's own route
    (audio_encoder/AudioMAE.py:356-366: torchaudio.functional.resample -> 16 kHz): the 16 kHz version of piano.wav that
    ``mel_piano_f16`` in audiomae_reference.npz was computed from, and a 2 s excerpt at the original 48 kHz with its
    reference-resampled counterpart."""
    import torchaudio
    from scipy.io import wavfile
    sr, raw = wavfile.read(os.path.join(ref_shims.REF_ROOT, "piano.wav"))
    assert sr == 48000 and raw.dtype == np.int16 and raw.ndim == 1
    wav = torch.from_numpy(raw.astype(np.float32) / 32768.0).unsqueeze(0)
    w16 = torchaudio.functional.resample(wav, orig_freq=sr, new_freq=16000)
    ex = raw[sr * 3: sr * 5]                                     # seconds 3..5
    ex16 = torchaudio.functional.resample(torch.from_numpy(ex.astype(np.float32) / 32768.0).unsqueeze(0), orig_freq=sr, new_freq=16000)
    np.savez_compressed(os.path.join(cases.GOLDEN_DIR, "frontend_reference.npz"), piano16k=w16[0].numpy(), piano48k_2s_i16=ex,
                        piano16k_2s=ex16[0].numpy(), sr=np.int64(sr))
    print(f"frontend: piano16k {tuple(w16.shape)}, excerpt {ex.shape} -> {tuple(ex16.shape)}")


def main():
    assert ref_shims.reference_available(), f"reference not found under {ref_shims.REF_ROOT}"
    os.makedirs(cases.GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    gen_real_weights()
    gen_ipattn()
    gen_audiomae()
    gen_frontend()
    for f in sorted(os.listdir(cases.GOLDEN_DIR)):
        print(f, os.path.getsize(os.path.join(cases.GOLDEN_DIR, f)))


if __name__ == "__main__":
    main()
