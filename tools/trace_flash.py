import torch
from ap_adapter_b200 import _lib
qkv = torch.randn(2, 1000, 768, device="cuda", dtype=torch.float16)
q, k, v = qkv.chunk(3, dim=-1)
for _ in range(2):
    _lib.flash_attention(q, k, v, 8)
torch.cuda.synchronize()
