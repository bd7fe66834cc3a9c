"""ORACLE (test infrastructure, not product): restatement of the denoising loop of
src/tryon_pipeline.py:1765-1866 and of diffusers==0.25.0 DDPMScheduler.set_timesteps/step (third-party, pinned by
environment.yaml:20, not vendored: restated from its published algorithm; "parity unpinned" for the scheduler).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import torch

from . import unet_ref as R


class DDPMRef:
    """diffusers DDPMScheduler: scaled_linear betas, epsilon prediction, fixed_small variance, leading spacing."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1,
                 rescale_betas_zero_snr=False):
        self.n = num_train_timesteps
        self.steps_offset = steps_offset
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        if rescale_betas_zero_snr:
            alphas_bar_sqrt = torch.cumprod(1.0 - betas, dim=0).sqrt()
            a0, aT = alphas_bar_sqrt[0].clone(), alphas_bar_sqrt[-1].clone()
            alphas_bar_sqrt = (alphas_bar_sqrt - aT) * (a0 / (a0 - aT))
            alphas_bar = alphas_bar_sqrt ** 2
            alphas = torch.cat([alphas_bar[0:1], alphas_bar[1:] / alphas_bar[:-1]])
            betas = 1 - alphas
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.one = torch.tensor(1.0)
        self.num_inference_steps = None

    def set_timesteps(self, num_inference_steps):
        self.num_inference_steps = num_inference_steps
        ratio = self.n // num_inference_steps
        ts = (torch.arange(0, num_inference_steps).double() * ratio).round().flip(0).long() + self.steps_offset
        self.timesteps = ts
        return ts

    def step(self, model_output, t, sample, generator=None, noise=None):
        t = int(t)
        prev_t = t - self.n // self.num_inference_steps
        alpha_prod_t = self.alphas_cumprod[t]
        alpha_prod_t_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        beta_prod_t = 1 - alpha_prod_t
        beta_prod_t_prev = 1 - alpha_prod_t_prev
        current_alpha_t = alpha_prod_t / alpha_prod_t_prev
        current_beta_t = 1 - current_alpha_t
        pred_original_sample = (sample - beta_prod_t ** 0.5 * model_output) / alpha_prod_t ** 0.5
        pred_original_sample_coeff = (alpha_prod_t_prev ** 0.5 * current_beta_t) / beta_prod_t
        current_sample_coeff = current_alpha_t ** 0.5 * beta_prod_t_prev / beta_prod_t
        pred_prev_sample = pred_original_sample_coeff * pred_original_sample + current_sample_coeff * sample
        variance = 0
        if t > 0:
            if noise is None:
                noise = torch.randn(model_output.shape, generator=generator, device=model_output.device,
                                    dtype=model_output.dtype)
            var = torch.clamp((1 - alpha_prod_t_prev) / (1 - alpha_prod_t) * current_beta_t, min=1e-20)
            variance = (var ** 0.5) * noise
        return pred_prev_sample + variance


def denoise_loop(sd_t, cfg_t, sd_g, cfg_g, inp, num_steps, guidance_scale=2.0, scheduler=None, noises=None,
                 max_steps=None, return_eps=False):
    """inp: dict with latents [B,4,h,w], mask/masked_image_latents/pose_latents [2B,..], cloth_latents [Bg,4,h,w],
    prompt_embeds [2B,77,X], add_text_embeds [2B,P], add_time_ids [2B,6], image_embeds [2B,16,X] (Resampler output),
    text_embeds_cloth [Bg,77,X]. noises: optional list of per-step variance-noise tensors (else seeded generator).
    Mirrors src/tryon_pipeline.py:1765-1823 line by line."""
    sch = scheduler or DDPMRef()
    timesteps = sch.set_timesteps(num_steps)
    latents = inp["latents"]
    eps_trace = []
    for i, t in enumerate(timesteps):
        if max_steps is not None and i >= max_steps:
            break
        latent_model_input = torch.cat([latents] * 2)                                            # :1769
        latent_model_input = torch.cat([latent_model_input, inp["mask"], inp["masked_image_latents"],
                                        inp["pose_latents"]], dim=1)                             # :1777
        tt = torch.as_tensor(int(t), device=latents.device)
        feats = R.unet_garment_forward(sd_g, cfg_g, inp["cloth_latents"], tt, inp["text_embeds_cloth"])  # :1787
        if feats[0].shape[0] != latents.shape[0]:
            feats = [f.expand(latents.shape[0], -1, -1) for f in feats]          # one shared garment (config 3)
        feats = [torch.cat([torch.zeros_like(d), d]) for d in feats]                              # :1796
        added = {"text_embeds": inp["add_text_embeds"], "time_ids": inp["add_time_ids"],
                 "image_embeds": inp["image_embeds"]}
        noise_pred = R.unet_tryon_forward(sd_t, cfg_t, latent_model_input, tt, inp["prompt_embeds"], added, feats)
        eps_trace.append(noise_pred)
        u, c = noise_pred.chunk(2)
        noise_pred = u + guidance_scale * (c - u)                                                 # :1815-1816
        latents = sch.step(noise_pred, t, latents, noise=None if noises is None else noises[i])   # :1823
    if return_eps:
        return latents, eps_trace
    return latents


def synth_loop_inputs(cfg_t, cfg_g, B, h, w, Bg=None, seed=0, device="cpu", dtype=torch.float32):
    """Seeded synthetic request: latents ~ N(0,1), rectangular mask, N(0,1) embeddings."""
    Bg = B if Bg is None else Bg
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    cross = cfg_t["cross_attention_dim"]
    pooled = cfg_t["projection_class_embeddings_input_dim"] - 6 * cfg_t["addition_time_embed_dim"]
    mask = torch.zeros(2 * B, 1, h, w)
    mask[:, :, h // 4: 3 * h // 4, w // 4: 3 * w // 4] = 1.0
    time_ids = torch.tensor([[h * 8.0, w * 8.0, 0.0, 0.0, h * 8.0, w * 8.0]]).repeat(2 * B, 1)
    d = dict(latents=r(B, 4, h, w), mask=mask, masked_image_latents=r(2 * B, 4, h, w) * 0.5,
             pose_latents=r(2 * B, 4, h, w) * 0.5, cloth_latents=r(Bg, 4, h, w) * 0.5,
             prompt_embeds=r(2 * B, 77, cross), add_text_embeds=r(2 * B, pooled), add_time_ids=time_ids,
             image_embeds=r(2 * B, cfg_t["ip_tokens"], cross), text_embeds_cloth=r(Bg, 77, cfg_g["cross_attention_dim"]))
    out = {}
    for k, v in d.items():
        out[k] = v.to(device=device, dtype=torch.float32 if k == "add_time_ids" and dtype == torch.float32 else dtype)
    return out
