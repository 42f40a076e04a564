"""ap_adapter_b200 — the AP-adapter hot path (decoupled attn2 processor + AudioMAE encoder) on B200 (sm_100a).

Public surface mirrors the reference:
  attention_processor.IPAttnProcessor2_0 / AttnProcessor2_0   (APadapter/ap_adapter/attention_processor.py)
  audiomae.AudioMAEConditionCTPoolRand / Vanilla_AudioMAE / extract_kaldi_fbank_feature   (audio_encoder/AudioMAE.py)
  unet.AudioLDM2UNet2DConditionModel (+ set_attn_processor / attn_processors)              (pipeline/modeling_audioldm2.py)
  pipeline.APAdapterDenoiser (audio conditioning block + CFG/DDIM loop)                    (pipeline/pipeline_audioldm2.py:919-1031)
All arithmetic of the hot path runs in libapx.so (include/apx.h); there is no CPU / eager fallback.
"""
from . import _lib  # noqa: F401
from .attention_processor import AttnProcessor2_0, IPAttnProcessor2_0  # noqa: F401

__version__ = "0.1.0"
