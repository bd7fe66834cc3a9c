"""Times the CLIP ViT-H image tower (hidden_states[-2], batch 2 = what one config-2 request batch encodes) on the engine's
kernels against the transformers module (fp16, SDPA) on the same GPU; device events, 10 iterations after 3 warm-ups.
Writes one JSON line to stdout."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timeit(fn, it=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def main():
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection, CLIPVisionConfig, CLIPVisionModelWithProjection
    import idm_vton_b200  # noqa: F401
    from idm_vton_b200 import lib as L
    from idm_vton_b200.clip import tower_for
    L.load()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    cfg = CLIPVisionConfig(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16,
                           patch_size=14, image_size=224, projection_dim=1024, hidden_act="gelu")
    m = CLIPVisionModelWithProjection(cfg).to(dev, torch.float16).eval()
    tower = tower_for(m)
    out = {}
    for B in (2, 8):
        x = torch.randn(B, 3, 224, 224, device=dev, dtype=torch.float16)
        with torch.no_grad():
            t_mod = timeit(lambda: m(x, output_hidden_states=True).hidden_states[-2])
        n0 = L.launch_count()
        tower.vision_hidden(x, -2)
        launches = L.launch_count() - n0
        t_eng = timeit(lambda: tower.vision_hidden(x, -2))
        # one forward captured in a CUDA graph (what a serving loop would replay)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            tower.vision_hidden(x, -2)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                y = tower.vision_hidden(x, -2)
        t_graph = timeit(g.replay)
        flops = B * 31 * (257 * (2 * 1280 * 3840 + 2 * 1280 * 1280 + 4 * 1280 * 5120) + 16 * 4 * 257 * 257 * 80)
        out[f"vit_h_B{B}"] = {"module_fp16_ms": round(t_mod, 3), "engine_eager_ms": round(t_eng, 3),
                              "engine_graph_ms": round(t_graph, 3), "launches": launches,
                              "engine_graph_tflops": round(flops / t_graph / 1e9, 1)}
    tc = CLIPTextConfig(vocab_size=49408, hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=20,
                        max_position_embeddings=77, hidden_act="gelu", projection_dim=1280, bos_token_id=49406,
                        eos_token_id=49407, pad_token_id=1)
    t = CLIPTextModelWithProjection(tc).to(dev, torch.float16).eval()
    tt = tower_for(t)
    ids = torch.randint(0, 49406, (2, 77), device=dev)
    ids[:, -1] = 49407
    with torch.no_grad():
        t_mod = timeit(lambda: t(ids, output_hidden_states=True))
    t_eng = timeit(lambda: tt.text_forward(ids))
    out["bigG_text_B2"] = {"module_fp16_ms": round(t_mod, 3), "engine_eager_ms": round(t_eng, 3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
