"""A/B of the VAE's fused pieces on one GPU: encoder pass of 6 images at 1024x768 + decoder pass of 2 latents, device events,
3 iterations after 1 warm-up, for the residual epilogue / attention kernels / fp16 GroupNorm -> convolution hand-off switched on one after the other. One JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import idm_vton_b200  # noqa: F401,E402
import idm_vton_b200.vae as V  # noqa: E402

torch.manual_seed(0)
vae = V.AutoencoderKL().to("cuda", torch.float32).eval()
x = torch.rand(6, 3, 1024, 768, device="cuda") * 2 - 1
z = torch.randn(2, 4, 128, 96, device="cuda")


def run():
    with torch.no_grad():
        m = vae.encode(x).latent_dist.mode()
        d = vae.decode(z, return_dict=False)[0]
    return m, d


out, keep = {}, {}
for name, fused, attn, f16 in (("aten", False, False, False), ("residual_epilogue", True, False, False),
                              ("fused", True, True, False), ("fused_f16_handoff", True, True, True)):
    V._ENGINE_FUSED, V._ATTN_FUSED, V._F16_ACT = fused, attn, f16
    keep[name] = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
    out[name + "_ms"] = round(e0.elapsed_time(e1) / 3, 2)
for name in ("residual_epilogue", "fused", "fused_f16_handoff"):
    out[name + "_vs_aten_maxdiff"] = [float((a - b).abs().max()) for a, b in zip(keep[name], keep["aten"])]
print(json.dumps(out))
