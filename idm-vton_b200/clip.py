"""CLIP towers on the engine's kernels.

The reference hands the pipeline three `transformers` modules: `image_encoder` (CLIPVisionModelWithProjection, ViT-H/14,
called twice per request at src/tryon_pipeline.py:468-470) and `text_encoder` / `text_encoder_2` (CLIPTextModel ViT-L and
CLIPTextModelWithProjection bigG, src/tryon_pipeline.py:592-612, called from inference.py:364,387). They stay the caller's
objects; `tower_for(module)` packs their weights once (same state-dict keys, nothing renamed) and runs the forward on
libb200vton: LayerNorm -> fused QKV GEMM (+bias) -> `b200vton_encoder_attention` (head dim 80 / 64, causal for text) ->
out-proj GEMM (+bias +residual) -> LayerNorm -> fc1 GEMM (+bias +GELU / quick-GELU) -> fc2 GEMM (+bias +residual), the patch
convolution as `b200vton_patchify` + GEMM (+position embedding as the residual term), the text embeddings as
`b200vton_token_embedding`. Rounding points are those of the fp16 module (fp16 after every linear / activation / residual
add, fp32 inside LayerNorm and softmax).

There is no fallback inside this file: an unsupported geometry makes `tower_for` return None and the pipeline then calls
the caller's module exactly as the reference does.
"""
import os
from types import SimpleNamespace

import torch

from . import lib as L

_ACTS = {"gelu": dict(gelu=True), "quick_gelu": dict(quick_gelu=True)}


def _cfg_get(cfg, name, default=None):
    return cfg.get(name, default) if isinstance(cfg, dict) else getattr(cfg, name, default)


class ClipTower:
    """One CLIP transformer tower (kind = "vision" | "text") with weights packed for the engine's kernels.

    state_dict: the module's own state dict (`vision_model.*` / `text_model.*`, `visual_projection.weight` /
    `text_projection.weight`); config: its transformers config (or a dict with the same fields)."""

    def __init__(self, state_dict, config, kind, device):
        assert kind in ("vision", "text")
        self.kind, self.device = kind, torch.device(device)
        g = lambda n, d=None: _cfg_get(config, n, d)
        self.C, self.heads, self.layers = int(g("hidden_size")), int(g("num_attention_heads")), int(g("num_hidden_layers"))
        self.D = self.C // self.heads
        self.eps = float(g("layer_norm_eps", 1e-5))
        act = g("hidden_act")
        if act not in _ACTS:
            raise ValueError(f"CLIP activation {act!r} is not built (gelu, quick_gelu)")
        self.act = _ACTS[act]
        if self.C % 64 or int(g("intermediate_size")) % 64 or self.D % 16 or not 16 <= self.D <= 96:
            raise ValueError(f"CLIP geometry hidden={self.C} heads={self.heads} is outside the kernels' range")
        root = "vision_model" if kind == "vision" else "text_model"
        f16 = lambda t: t.detach().to(self.device, torch.float16).contiguous()
        sd = state_dict
        self.blocks = []
        for i in range(self.layers):
            p = f"{root}.encoder.layers.{i}"
            a = f"{p}.self_attn"
            self.blocks.append(SimpleNamespace(
                ln1=(f16(sd[f"{p}.layer_norm1.weight"]), f16(sd[f"{p}.layer_norm1.bias"])),
                ln2=(f16(sd[f"{p}.layer_norm2.weight"]), f16(sd[f"{p}.layer_norm2.bias"])),
                wqkv=f16(torch.cat([sd[f"{a}.q_proj.weight"], sd[f"{a}.k_proj.weight"], sd[f"{a}.v_proj.weight"]], 0)),
                bqkv=f16(torch.cat([sd[f"{a}.q_proj.bias"], sd[f"{a}.k_proj.bias"], sd[f"{a}.v_proj.bias"]], 0)),
                wo=f16(sd[f"{a}.out_proj.weight"]), bo=f16(sd[f"{a}.out_proj.bias"]),
                w1=f16(sd[f"{p}.mlp.fc1.weight"]), b1=f16(sd[f"{p}.mlp.fc1.bias"]),
                w2=f16(sd[f"{p}.mlp.fc2.weight"]), b2=f16(sd[f"{p}.mlp.fc2.bias"])))
        if kind == "vision":
            w = sd[f"{root}.embeddings.patch_embedding.weight"]                  # [C, 3, P, P], no bias
            self.P, self.K = int(w.shape[-1]), int(w[0].numel())
            self.Kp = (self.K + 63) // 64 * 64
            wp = torch.zeros(self.C, self.Kp, dtype=torch.float16, device=self.device)
            wp[:, :self.K] = f16(w).reshape(self.C, self.K)
            self.w_patch = wp
            pos = f16(sd[f"{root}.embeddings.position_embedding.weight"])         # [1 + G, C]
            cls = f16(sd[f"{root}.embeddings.class_embedding"])
            self.pos_patches = pos[1:].contiguous()
            self.cls_pos0 = (cls + pos[0]).contiguous()                           # fp16(cls + pos[0]): a constant of the weights
            self.pre_ln = (f16(sd[f"{root}.pre_layrnorm.weight"]), f16(sd[f"{root}.pre_layrnorm.bias"]))
            self.post_ln = (f16(sd[f"{root}.post_layernorm.weight"]), f16(sd[f"{root}.post_layernorm.bias"]))
            self.proj = f16(sd["visual_projection.weight"]) if "visual_projection.weight" in sd else None
        else:
            self.tok = f16(sd[f"{root}.embeddings.token_embedding.weight"])
            self.pos = f16(sd[f"{root}.embeddings.position_embedding.weight"])
            self.final_ln = (f16(sd[f"{root}.final_layer_norm.weight"]), f16(sd[f"{root}.final_layer_norm.bias"]))
            self.proj = f16(sd["text_projection.weight"]) if "text_projection.weight" in sd else None
            self.eos_token_id = g("eos_token_id", 2)

    # ---------------------------------------------------------------------------------------------
    def _layer(self, x, B, T, blk):
        C = self.C
        h = L.layernorm(x, *blk.ln1, eps=self.eps)
        qkv = L.gemm(h, blk.wqkv, bias=blk.bqkv).view(B, T, 3 * C)
        o = L.encoder_attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], self.heads, self.D,
                                causal=self.kind == "text")
        x = L.gemm(o.view(B * T, C), blk.wo, bias=blk.bo, residual=x)
        h = L.layernorm(x, *blk.ln2, eps=self.eps)
        h = L.gemm(h, blk.w1, bias=blk.b1, **self.act)
        return L.gemm(h, blk.w2, bias=blk.b2, residual=x)

    def _run(self, x, B, T, n_layers, collect):
        states = [x.view(B, T, self.C)] if collect else None
        for blk in self.blocks[:n_layers]:
            x = self._layer(x, B, T, blk)
            if collect:
                states.append(x.view(B, T, self.C))
        return x, states

    def _project(self, rows, ln):
        """[B, C] pooled rows -> LayerNorm -> projection (no bias); skinny linears take 16 rows per launch."""
        pooled = L.layernorm(rows.contiguous(), *ln, eps=self.eps)
        if self.proj is None:
            return pooled, None
        emb = torch.cat([L.skinny_linear(pooled[i:i + 16], self.proj) for i in range(0, pooled.shape[0], 16)])
        return pooled, emb

    # ---------------------------------------------------------------------------------------------
    def vision_embed(self, pixel_values):
        """CLIPVisionEmbeddings + pre_layrnorm: [B,3,H,W] -> [B*(1+G), C] (= hidden_states[0])."""
        assert self.kind == "vision"
        x = pixel_values.to(self.device, torch.float16).contiguous()
        B = x.shape[0]
        G = (x.shape[2] // self.P) * (x.shape[3] // self.P)
        if G != self.pos_patches.shape[0]:
            raise ValueError(f"image of {x.shape[2]}x{x.shape[3]} px gives {G} patches, the position table has "
                             f"{self.pos_patches.shape[0]} (interpolate_pos_encoding is not built)")
        T = G + 1
        a = L.patchify(x, self.P, self.Kp)
        hs = torch.empty(B, T, self.C, dtype=torch.float16, device=self.device)
        for b in range(B):   # fp16(fp16(conv) + pos) straight into rows 1.. of the token matrix
            L.gemm(a[b * G:(b + 1) * G], self.w_patch, residual=self.pos_patches, out=hs[b, 1:])
        hs[:, 0] = self.cls_pos0
        return L.layernorm(hs.view(B * T, self.C), *self.pre_ln, eps=self.eps), B, T

    def vision_hidden(self, pixel_values, index=-2):
        """hidden_states[index] of the module (index counted like the module's tuple of 1 + layers entries), running only
        the layers it needs: index = -2 (what src/tryon_pipeline.py:468 reads) skips the last block."""
        x, B, T = self.vision_embed(pixel_values)
        n = index if index >= 0 else self.layers + 1 + index
        if not 0 <= n <= self.layers:
            raise IndexError(f"hidden state {index} of a {self.layers}-layer tower")
        x, _ = self._run(x, B, T, n, False)
        return x.view(B, T, self.C)

    def vision_forward(self, pixel_values, output_hidden_states=False):
        """The module's forward: last_hidden_state, pooler_output (post_layernorm of the class token), image_embeds and,
        on request, all 1 + layers hidden states."""
        x, B, T = self.vision_embed(pixel_values)
        x, states = self._run(x, B, T, self.layers, output_hidden_states)
        last = x.view(B, T, self.C)
        pooled, emb = self._project(last[:, 0], self.post_ln)
        return SimpleNamespace(last_hidden_state=last, pooler_output=pooled, image_embeds=emb,
                               hidden_states=tuple(states) if states is not None else None)

    def text_forward(self, input_ids, output_hidden_states=True):
        """The module's forward for [B,T] token ids: hidden_states (embeddings + every block, no final LayerNorm),
        last_hidden_state (after final_layer_norm), pooler_output (its EOS row) and text_embeds (projected)."""
        assert self.kind == "text"
        ids = input_ids.to(self.device, torch.int64).contiguous()
        B, T = ids.shape
        if T > self.pos.shape[0]:
            raise ValueError(f"{T} tokens exceed the {self.pos.shape[0]}-entry position table")
        x = L.token_embedding(ids.view(-1), self.tok, self.pos, T)
        x, states = self._run(x, B, T, self.layers, output_hidden_states)
        last = L.layernorm(x, *self.final_ln, eps=self.eps).view(B, T, self.C)
        if self.eos_token_id == 2:      # legacy configs: the EOS token has the highest id
            eos = ids.argmax(dim=-1)
        else:
            eos = (ids == self.eos_token_id).int().argmax(dim=-1)
        pooled = last[torch.arange(B, device=self.device), eos]
        emb = None
        if self.proj is not None:
            emb = torch.cat([L.skinny_linear(pooled[i:i + 16].contiguous(), self.proj) for i in range(0, B, 16)])
        return SimpleNamespace(last_hidden_state=last, pooler_output=pooled, text_embeds=emb,
                               hidden_states=tuple(states) if states is not None else None)


# ------------------------------------------------------------------------------------------------
_towers = {}


def _version(module):
    return tuple((p.data_ptr(), p._version) for p in module.parameters())


def tower_for(module):
    """The engine twin of a transformers CLIP vision / text module, packed once per weight version; None when the module
    is not one the kernels cover (not on a GPU, not fp16, unsupported geometry or activation) or B200VTON_CLIP=0."""
    if module is None or os.environ.get("B200VTON_CLIP", "1") == "0":
        return None
    cfg = getattr(module, "config", None)
    name = type(module).__name__
    kind = "vision" if "Vision" in name else ("text" if "Text" in name else None)
    if cfg is None or kind is None or not name.startswith("CLIP"):
        return None
    try:
        p0 = next(module.parameters())
    except StopIteration:
        return None
    if not p0.is_cuda or p0.dtype != torch.float16:
        return None
    key = id(module)
    ver = _version(module)
    hit = _towers.get(key)
    if hit is not None and hit[0] == ver and hit[2] is module:
        return hit[1]
    try:
        tower = ClipTower(module.state_dict(), cfg, kind, p0.device)
    except (ValueError, KeyError):
        _towers[key] = (ver, None, module)
        return None
    _towers[key] = (ver, tower, module)
    return tower
