"""Diagnostic: is the serving scenario of tests/test_seams_gpu.py::test_serving_front_end_garment_batching_and_kv_cache
bit-reproducible, and if not, which stage differs? Records, per pipeline call, the CLIP tokens, the Resampler output, the
final latents and the image; optionally poisons the caching allocator's free memory with NaN between runs (a kernel that
reads memory it did not write then shows up as NaN instead of as a 1e-4 wobble). One JSON line per scenario."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def poison(gb=8):
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    blocks = [torch.full((gb << 29,), float("nan"), dtype=torch.float16, device="cuda")]
    small = [torch.full((1 << 20,), float("nan"), dtype=torch.float16, device="cuda") for _ in range(64)]
    small += [torch.full((1 << 14,), float("nan"), dtype=torch.float16, device="cuda") for _ in range(256)]
    torch.cuda.synchronize()
    del blocks, small


def scenario(do_poison):
    from oracle import make_golden_pipeline as MG
    from oracle import unet_ref as R
    from idm_vton_b200 import unet as U
    from idm_vton_b200.pipeline import StableDiffusionXLInpaintPipeline
    from idm_vton_b200.scheduler import DDPMScheduler
    from idm_vton_b200.serving import TryOnRequest, TryOnServer
    dev, f16 = "cuda", torch.float16
    cfg_t, cfg_g = R.tiny_config("tryon"), R.tiny_config("garment")
    net_t = U.UNet2DConditionModel(cfg_t, R.make_state_dict(cfg_t, seed=11)).to(dev, f16)
    net_g = U.UNet2DConditionModelGarment(cfg_g, R.make_state_dict(cfg_g, seed=22)).to(dev, f16)
    pipe = StableDiffusionXLInpaintPipeline(
        vae=MG.make_vae().to(dev, f16), text_encoder=None, text_encoder_2=None, tokenizer=None, tokenizer_2=None,
        unet=net_t, unet_encoder=net_g, scheduler=DDPMScheduler(),
        image_encoder=MG.make_image_encoder(cfg_t["resampler"]["embedding_dim"]).to(dev, f16))
    rec = []
    cur = {}
    enc0, hid0 = pipe.encode_image, pipe.unet.encoder_hid_proj.forward

    def enc(*a, **k):
        out = enc0(*a, **k)
        cur["clip_cond"], cur["clip_uncond"] = out[0].clone(), out[1].clone()
        return out

    def hid(x):
        y = hid0(x)
        cur["resampler"] = y.clone()
        return y
    pipe.encode_image = enc
    pipe.unet.encoder_hid_proj.forward = hid
    if "--eager" in sys.argv:
        pipe.use_cuda_graph = False
    from idm_vton_b200 import denoise as DN
    names = ("latents", "mask", "masked_image_latents", "pose_latents", "cloth_latents", "prompt_embeds", "add_text_embeds",
             "add_time_ids", "image_embeds", "text_embeds_cloth")
    prep0, tables0, step0 = DN.TryOnDenoiser.prepare, DN.TryOnDenoiser.set_step_tables, DN.TryOnDenoiser.step

    def prep(self, *a, **k):
        for n, v in zip(names, a):
            cur["in." + n] = v.detach().clone()
        r = prep0(self, *a, **k)
        cur["prep.x_t"] = self.x_t.clone()
        cur["prep.aug"] = self.aug.clone()
        cur["prep.ctx_t"] = torch.cat([t.reshape(-1).float() for pair in self.ctx_t for t in pair if t is not None])
        return r

    def tables(self, *a, **k):
        r = tables0(self, *a, **k)
        cur["gkv"] = torch.cat([g.reshape(-1)[:: max(1, g.numel() // 65536)].float() for g in self.gkv_all])
        cur["gkv_sum"] = torch.stack([g.double().sum() for g in self.gkv_all])
        return r

    def step(self, i, noise=None, use_graph=True):
        if noise is not None:
            cur[f"noise{i}"] = noise.clone()
        r = step0(self, i, noise, use_graph)
        cur[f"step{i}.latents"] = r.clone()
        cur[f"step{i}.eps"] = self.eps.clone()
        return r
    DN.TryOnDenoiser.prepare, DN.TryOnDenoiser.set_step_tables, DN.TryOnDenoiser.step = prep, tables, step

    def req(gid, seed):
        i = MG.make_call_inputs(cfg_t, B=1, seed=seed)
        gi = MG.make_call_inputs(cfg_t, B=1, seed=1000 + {"A": 1, "B": 2}[gid])
        return TryOnRequest(garment_id=gid, image=i["image"][0], mask_image=i["mask_image"][0], pose_img=i["pose_img"][0],
                            prompt_embeds=i["prompt_embeds"][0], negative_prompt_embeds=i["negative_prompt_embeds"][0],
                            pooled_prompt_embeds=i["pooled_prompt_embeds"][0],
                            negative_pooled_prompt_embeds=i["negative_pooled_prompt_embeds"][0], cloth=gi["cloth"][0],
                            ip_adapter_image=gi["ip_adapter_image"][0], text_embeds_cloth=gi["text_embeds_cloth"][0])

    srv = TryOnServer(pipe, height=MG.H, width=MG.W, num_inference_steps=3, guidance_scale=2.0, max_batch=4, seed=7)

    def run(reqs):
        for r in reqs:
            srv.submit(r)
        outs = []
        while srv.queue:
            cur.clear()
            o = srv.step()
            cur["latents"] = pipe._last_latents.clone()
            cur["image"] = torch.stack([o[k] for k in sorted(o)]).clone()
            outs.append(dict(cur))
        return outs
    r1 = run([req("A", 1), req("A", 2), req("B", 3)])
    if do_poison:
        poison()
    r2 = run([req("A", 1), req("A", 2)])
    if do_poison:
        poison()
    r3 = run([req("A", 1), req("A", 2)])
    res = {}
    for name, other in (("run2", r2[0]), ("run3", r3[0])):
        for k in sorted(r1[0]):
            if k not in other:
                res[f"{name}.{k}"] = "missing"
                continue
            a, b = r1[0][k].float(), other[k].float()
            d = (a - b).abs()
            if not torch.equal(a, b):      # only what differs is reported
                res[f"{name}.{k}"] = f"max {d.max().item():.3e} n_diff {int((d > 0).sum())}/{d.numel()} nan {int(torch.isnan(b).sum())}"
    res["keys"] = len(r1[0])
    res["cache_hits"] = pipe.garment_cache.hits
    return res


if __name__ == "__main__":
    print(json.dumps({"clip_env": os.environ.get("B200VTON_CLIP", "1"), "argv": sys.argv[1:],
                      "pdl_graph": os.environ.get("B200VTON_PDL_GRAPH", "1"), "poison": "--poison" in sys.argv,
                      **scenario("--poison" in sys.argv)}))
