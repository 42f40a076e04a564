"""apx_flash_attention (csrc/apx_flash_attn.cu): the plain softmax attention of AttnProcessor2_0 (reference
APadapter/ap_adapter/attention_processor.py:267-274) at the UNet's self-attention sites, against fp32 torch on the same 16-bit
operands."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import record, rel_err

pytestmark = pytest.mark.gpu


def _ref(q, k, v, heads):
    b, s, c = q.shape
    d = c // heads
    qh, kh, vh = (t.float().view(b, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    return F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(b, s, c)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
@pytest.mark.parametrize("b,s,l,c,heads", [(2, 1000, 1000, 256, 8), (2, 252, 252, 384, 8), (2, 64, 64, 640, 8), (1, 130, 130, 128, 8),
                                           (2, 300, 77, 256, 8), (1, 1, 129, 64, 4), (3, 128, 384, 512, 8), (16, 1000, 1000, 256, 8)],
                         ids=["self256", "self384", "self640", "d16", "cross77", "one_query", "three_blocks", "batch8"])
def test_flash_attention_matches_fp32_sdpa(b, s, l, c, heads, dtype):
    from ap_adapter_b200 import _lib
    g = torch.Generator().manual_seed(s + c)
    q, k, v = (torch.randn(b, n, c, generator=g).to(dtype).cuda() for n in (s, l, l))
    q = q * 1.5                                   # sharper rows: the running maximum moves between key blocks
    n0 = _lib.launch_count()
    o = _lib.flash_attention(q, k, v, heads)
    assert _lib.launch_count() - n0 == 1 and o.shape == (b, s, c)
    err = rel_err(o, _ref(q, k, v, heads))
    record(f"flash/{b}x{s}x{l}x{c}", dtype=str(dtype), err=err)
    assert err < (1.5e-3 if dtype == torch.float16 else 1.2e-2)
    assert torch.equal(o, _lib.flash_attention(q, k, v, heads))


def test_flash_attention_on_the_slices_of_a_fused_qkv_projection():
    """q / k / v as column slices of one [B, S, 3C] tensor (what the fused QKV projection produces): strided views, no copies."""
    from ap_adapter_b200 import _lib
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(2, 500, 3 * 256, generator=g).half().cuda()
    q, k, v = qkv.chunk(3, dim=-1)
    o = _lib.flash_attention(q, k, v, 8)
    assert rel_err(o, _ref(q, k, v, 8)) < 1.5e-3


def test_flash_attention_rejects_unsupported_head_sizes():
    from ap_adapter_b200 import _lib
    x = torch.randn(1, 64, 320, device="cuda", dtype=torch.float16)      # d = 40
    assert not _lib.flash_attention_ok(320, 8)
    with pytest.raises(_lib.ApxError, match="head dim"):
        _lib.flash_attention(x, x, x, 8)
