"""CPU oracle for the AP-adapter hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

A plain fp32 restatement (torch on CPU) of the reference algorithms the CUDA kernels replace:

* ``ipattn_oracle``   — ``IPAttnProcessor2_0.__call__`` / ``AttnProcessor2_0.__call__``
  (reference ``APadapter/ap_adapter/attention_processor.py:347-470`` and ``:214-294``)
* ``audiomae_oracle`` — ``forward_encoder_no_random_mask_no_average`` + ``pool``
  (reference ``audio_encoder/models_mae.py:19-44,136-151,548-570,689-701``, ``audio_encoder/AudioMAE.py:148-212``;
  timm ``Block`` semantics restated from SURVEY.md appendix A — timm itself is not vendored in the reference)

Parity pinning: the reference ships no tests or golden vectors, so the oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF: ``oracle/gen_golden.py`` imports the reference's own modules from
``/root/reference`` (with ``sys.modules`` shims for the absent ``timm``/``librosa``/``matplotlib``), runs them on
seeded inputs and commits the results under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks this
restatement against those fixtures; the GPU parity tests check the CUDA path against this restatement.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import this package.  The product package (``ap_adapter_b200``) never does.
"""
