"""Why does the device-timed loop of bench.py read slower than its end-to-end loop?  Times the same captured step graph four ways
(200 replays each, twice, interleaved): replay only; replay + the bench's three device-side input copies; the same with the
copies batched into one foreach copy; the end-to-end loop (pinned-host inputs, result to host, sync per step).
Run on the GPU box: PYTHONPATH=. python tools/time_step_loops.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ap_adapter_b200.attention_processor import IPAttnProcessor2_0  # noqa: E402
from ap_adapter_b200.pipeline import APAdapterDenoiser  # noqa: E402
from ap_adapter_b200.unet import AudioLDM2UNet2DConditionModel, build_attn_processors  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    unet = AudioLDM2UNet2DConditionModel().eval().half().to(dev)
    unet.set_attn_processor(dict(build_attn_processors(unet, IPAttnProcessor2_0, ap_scale=bench.AP_SCALE, device=dev, dtype=torch.float16)))
    den = APAdapterDenoiser(unet)
    lat0, gpe, pe, mask = bench.clip_inputs(range(1), dev)
    den.scheduler.set_timesteps(bench.INFER_STEPS, device=dev)
    ts, coefs = den.scheduler.timesteps, den.scheduler.coefficients(dev)
    n = 200
    with torch.no_grad():
        g = den._step_graph(lat0, gpe, pe, mask, bench.GUIDANCE)
        h_lat, h_t, h_coef = lat0.cpu().pin_memory(), ts.cpu().pin_memory(), coefs.cpu().pin_memory()

        def replay_only(i):
            g["graph"].replay()

        def bench_step(i):
            g["t"].copy_(ts[i % bench.INFER_STEPS])
            g["coef"].copy_(coefs[i % bench.INFER_STEPS])
            g["graph"].replay()
            g["latents"].copy_(g["out"])

        def e2e_step(i):
            g["latents"].copy_(h_lat, non_blocking=True)
            g["t"].copy_(h_t[i % bench.INFER_STEPS], non_blocking=True)
            g["coef"].copy_(h_coef[i % bench.INFER_STEPS], non_blocking=True)
            g["graph"].replay()
            h_lat.copy_(g["out"], non_blocking=True)
            torch.cuda.current_stream().synchronize()

        def replay_sync(i):
            g["graph"].replay()
            torch.cuda.current_stream().synchronize()

        def timed(fn):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g["latents"].copy_(lat0)
            h_lat.copy_(lat0)
            torch.cuda.synchronize()
            e0.record()
            for i in range(n):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n

        for i in range(100):
            bench_step(i)
        torch.cuda.synchronize()
        for rep in range(3):
            print(f"rep {rep}: " + "  ".join(f"{name} {timed(fn):.3f} ms" for name, fn in
                                             (("replay_only", replay_only), ("bench_step", bench_step), ("replay+sync", replay_sync), ("e2e", e2e_step))), flush=True)


if __name__ == "__main__":
    main()
