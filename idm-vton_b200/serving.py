"""Request front-end for the try-on engine: replaces the reference's DataLoader loop
(inference.py:309-341, 337-419: fixed batches in dataset order, the garment UNet re-run for every sample of every batch)
with batching BY GARMENT and reuse of garment work across requests.

  * requests that wear the same garment are batched together: the garment UNet then runs at batch 1 and its K/V of every
    denoise step are indexed by all persons of the batch (BASELINE config 3);
  * each garment is VAE-encoded ONCE (one posterior sample per garment instead of one per call) and the hoisted garment
    K/V of all denoise steps are kept in an LRU cache (denoise.GarmentKVCache), so a garment seen before costs a few
    device-to-device copies instead of `num_inference_steps` garment-UNet passes.
Everything per person (noise, masks, pose, prompts) is exactly what `StableDiffusionXLInpaintPipeline.__call__` does; the
only semantic difference from calling the pipeline per DataLoader batch is the once-per-garment posterior sample, which is
why the cache is a property of this front-end and not of the pipeline (off unless a server installs it).
Single-threaded by design, like the reference pipeline (one CUDA stream, one denoiser / graph per process).
"""
import collections
import dataclasses
from typing import Any, Hashable, Optional

import torch

from .denoise import GarmentKVCache


@dataclasses.dataclass
class TryOnRequest:
    """One person x one garment. Tensors follow the keyword set of inference.py:397-414 for a batch of 1."""
    garment_id: Hashable
    image: torch.Tensor                 # [3,H,W] in [0,1]
    mask_image: torch.Tensor            # [1,H,W]
    pose_img: torch.Tensor              # [3,H,W] in [-1,1]
    prompt_embeds: torch.Tensor         # [77,2048]
    negative_prompt_embeds: torch.Tensor
    pooled_prompt_embeds: torch.Tensor  # [1280]
    negative_pooled_prompt_embeds: torch.Tensor
    # garment side (needed the first time a garment_id is seen; ignored afterwards)
    cloth: Optional[torch.Tensor] = None             # [3,H,W] in [-1,1]
    ip_adapter_image: Optional[torch.Tensor] = None  # [3,224,224] CLIP-preprocessed garment image
    text_embeds_cloth: Optional[torch.Tensor] = None  # [77,2048]
    ticket: Any = None


class TryOnServer:
    def __init__(self, pipe, height=1024, width=768, num_inference_steps=30, guidance_scale=2.0, max_batch=8, seed=None,
                 garment_cache_bytes=40 << 30, output_type="pt"):
        self.pipe = pipe
        self.height, self.width = height, width
        self.num_inference_steps, self.guidance_scale = num_inference_steps, guidance_scale
        self.max_batch = max_batch
        self.seed = seed
        self.output_type = output_type
        self.queue = collections.OrderedDict()      # garment_id -> deque of requests (arrival order inside a garment)
        self.garments = {}                          # garment_id -> dict(latents, ip_adapter_image, text_embeds_cloth)
        self._next_ticket = 0
        self.stats = collections.Counter()
        if garment_cache_bytes:
            pipe.garment_cache = GarmentKVCache(garment_cache_bytes)

    # ---------------------------------------------------------------------------------------------
    def submit(self, req: TryOnRequest):
        if req.garment_id not in self.garments and req.garment_id not in self.queue and \
                (req.cloth is None or req.ip_adapter_image is None or req.text_embeds_cloth is None):
            raise ValueError(f"garment {req.garment_id!r} is new: cloth, ip_adapter_image and text_embeds_cloth are required")
        req.ticket = self._next_ticket
        self._next_ticket += 1
        self.queue.setdefault(req.garment_id, collections.deque()).append(req)
        return req.ticket

    def pending(self):
        return sum(len(q) for q in self.queue.values())

    def _next_batch(self):
        """Oldest waiting request decides the garment; up to max_batch requests of that garment go together."""
        gid = min(self.queue, key=lambda g: self.queue[g][0].ticket)
        q = self.queue[gid]
        batch = [q.popleft() for _ in range(min(self.max_batch, len(q)))]
        if not q:
            del self.queue[gid]
        return gid, batch

    def _garment(self, gid, batch, device, dtype):
        g = self.garments.get(gid)
        if g is None:
            src = next(r for r in batch if r.cloth is not None)
            cloth = src.cloth[None].to(device=device, dtype=dtype)
            # ONE posterior sample per garment, from a generator of its own: the per-request generator handed to the
            # pipeline must see the same stream whether or not the garment was already known
            gen = torch.Generator(device).manual_seed(self.seed) if self.seed is not None else None
            latents = self.pipe._encode_vae_image(cloth, generator=gen)
            g = dict(latents=latents, ip_adapter_image=src.ip_adapter_image[None].to(device),
                     text_embeds_cloth=src.text_embeds_cloth[None].to(device=device, dtype=dtype))
            self.garments[gid] = g
            self.stats["garments_encoded"] += 1
        return g

    @torch.no_grad()
    def step(self):
        """Runs ONE batch; returns {ticket: image}."""
        if not self.queue:
            return {}
        gid, batch = self._next_batch()
        pipe = self.pipe
        device = pipe._execution_device
        dtype = pipe.unet.dtype
        gen = torch.Generator(device).manual_seed(self.seed) if self.seed is not None else None
        g = self._garment(gid, batch, device, dtype)
        stack = lambda name, dt=None: torch.stack([getattr(r, name) for r in batch]).to(device=device, dtype=dt)  # noqa: E731
        # The reference draws the pose latents' posterior sample from the GLOBAL generator (src/tryon_pipeline.py:1646 passes no
        # generator), so a seeded server would still not be reproducible: with a seed, the global CPU / device generators are
        # forked around the call and seeded too (their state outside the call is untouched).
        import contextlib
        dev_idx = [torch.device(device).index or 0] if torch.device(device).type == "cuda" else []
        with (torch.random.fork_rng(devices=dev_idx) if self.seed is not None else contextlib.nullcontext()):
            if self.seed is not None:
                torch.manual_seed(self.seed)
            images = pipe(prompt_embeds=stack("prompt_embeds", dtype), negative_prompt_embeds=stack("negative_prompt_embeds", dtype),
                          pooled_prompt_embeds=stack("pooled_prompt_embeds", dtype),
                          negative_pooled_prompt_embeds=stack("negative_pooled_prompt_embeds", dtype),
                          num_inference_steps=self.num_inference_steps, generator=gen, strength=1.0,
                          pose_img=stack("pose_img", dtype), text_embeds_cloth=g["text_embeds_cloth"], cloth=g["latents"],
                          mask_image=stack("mask_image"), image=stack("image"), height=self.height, width=self.width,
                          ip_adapter_image=g["ip_adapter_image"], guidance_scale=self.guidance_scale,
                          output_type=self.output_type, garment_keys=[gid])[0]
        self.stats["batches"] += 1
        self.stats["images"] += len(batch)
        return {r.ticket: images[i] for i, r in enumerate(batch)}

    def run(self):
        """Drains the queue; returns {ticket: image}."""
        out = {}
        while self.queue:
            out.update(self.step())
        return out
