"""apx_groupnorm_nhwc at the UNet's shapes (BASELINE configs[1], Bc = 2), CUDA-graph replay of 20 back-to-back calls (test infrastructure).
APX_GN_V1=1 selects the first form of the kernel.    PYTHONPATH=.:tools python tools/time_norm.py"""
import torch

from ap_adapter_b200 import _lib
from time_gemm_step import graph_time

SHAPES = [  # (H, W, C): the 10 s latent at the four levels, block widths and the concatenated widths of the up path
    (250, 16, 128), (250, 16, 256), (250, 16, 384), (125, 8, 128), (125, 8, 256), (125, 8, 384), (125, 8, 512), (125, 8, 640),
    (63, 4, 256), (63, 4, 384), (63, 4, 640), (63, 4, 768), (63, 4, 1024), (32, 2, 384), (32, 2, 640), (32, 2, 1280),
]


def main():
    torch.manual_seed(0)
    tot = 0.0
    for h, w, c in SHAPES:
        x = torch.randn(2, c, h, w, device="cuda", dtype=torch.float16).contiguous(memory_format=torch.channels_last)
        gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
        cb, t = torch.randn(c, device="cuda") * 0.1, torch.randn(2, c, device="cuda", dtype=torch.float16)
        ta = graph_time(lambda: _lib.groupnorm_nhwc(x, gamma, beta, 32, 1e-5, True, cb, t))
        gn = torch.nn.GroupNorm(32, c, eps=1e-5).cuda().half()

        def ref():
            return torch.nn.functional.silu(gn(x + cb.half()[None, :, None, None] + t[:, :, None, None]))

        tt = graph_time(ref)
        tot += ta
        print(f"B=2 HW={h * w:5d} C={c:5d}: apx {ta:6.2f} us   torch (3 adds + GroupNorm + SiLU) {tt:6.2f} us   {2 * h * w * c * 4 / ta / 1e3:7.1f} GB/s")
    print(f"sum over shapes: {tot:.1f} us")


if __name__ == "__main__":
    main()
