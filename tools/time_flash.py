"""apx_flash_attention against torch's SDPA (cuDNN) at the UNet's self-attention shapes, CUDA-graph replay of 20 calls (test infrastructure)."""
import torch
import torch.nn.functional as F
from ap_adapter_b200 import _lib
from time_gemm_step import graph_time

for b, s, c, h in [(2, 1000, 256, 8), (2, 252, 384, 8), (2, 64, 640, 8), (16, 1000, 256, 8), (16, 252, 384, 8)]:
    qkv = torch.randn(b, s, 3 * c, device="cuda", dtype=torch.float16)
    q, k, v = qkv.chunk(3, dim=-1)
    d = c // h

    def ref():
        qh, kh, vh = (t.view(b, s, h, d).transpose(1, 2) for t in (q, k, v))
        return F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(b, s, c)

    ta, tt = graph_time(lambda: _lib.flash_attention(q, k, v, h)), graph_time(ref)
    fl = 4.0 * b * h * s * s * d
    print(f"B={b} S={s} C={c}: apx {ta:7.2f} us ({fl / ta / 1e6:6.1f} TFLOP/s)   torch SDPA {tt:7.2f} us")
