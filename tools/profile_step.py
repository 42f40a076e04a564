"""Where one eager UNet denoise step spends its GPU time, by (kernel-launching op, input shapes, python call site).
Run on the GPU box: PYTHONPATH=. python tools/profile_step.py > gpurun_out/step_profile.txt"""
import collections
import torch
from torch.profiler import ProfilerActivity, profile

from ap_adapter_b200.attention_processor import AttnProcessor2_0, IPAttnProcessor2_0
from ap_adapter_b200.unet import AudioLDM2UNet2DConditionModel, build_attn_processors


def main():
    torch.manual_seed(0)
    unet = AudioLDM2UNet2DConditionModel().eval().half().cuda()
    unet.set_attn_processor(build_attn_processors(unet, IPAttnProcessor2_0, AttnProcessor2_0, ap_scale=0.5, device="cuda", dtype=torch.float16))
    x = torch.randn(2, 8, 250, 16, device="cuda", dtype=torch.float16)
    kw = dict(encoder_hidden_states=torch.randn(2, 136, 768, device="cuda", dtype=torch.float16),
              encoder_hidden_states_1=torch.randn(2, 16, 1024, device="cuda", dtype=torch.float16),
              encoder_attention_mask_1=torch.ones(2, 16, device="cuda"))
    t = torch.tensor(500, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            unet(x, t, **kw)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
            unet(x, t, **kw)
            torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        dt = ev.self_device_time_total
        if dt <= 0:
            continue
        site = ""
        for fr in ev.stack:
            if "ap_adapter_b200" in fr:
                site = fr.split("ap_adapter_b200/")[-1]
                break
        shapes = str(ev.input_shapes)[:70]
        a = agg[(ev.name, shapes, site)]
        a[0] += dt
        a[1] += 1
    total = sum(v[0] for v in agg.values())
    print(f"total device time {total:.0f} us over {sum(v[1] for v in agg.values())} kernel-launching ops")
    for (name, shapes, site), (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:70]:
        print(f"{us:9.1f} us {100 * us / total:5.1f}% n={n:4d}  {name[:38]:38s} {shapes:70s} {site}")


if __name__ == "__main__":
    main()
