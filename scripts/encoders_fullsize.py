"""HunyuanVideo's prompt stack and first-frame VAE encode at their TRUE sizes on one B200 (seeded weights, native kernels):
LLaVA-Llama-3-8B (CLIP-ViT-L/14-336 tower + projector + 32 Llama layers, 934 expanded tokens), CLIP-L text (77 tokens), and
AutoencoderKLHunyuanVideo.encode of one 720 x 1280 frame (hy:576-581).  Writes gpurun_out/r02_hunyuan_aux_fullsize.json."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alg_b200 import encoders, llava  # noqa: E402
from alg_b200.vae_hunyuan import AutoencoderKLHunyuanVideo  # noqa: E402


def timed(fn, n=2):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return out, a.elapsed_time(b) / n


def main():
    dev = "cuda"
    res = {}
    t0 = time.time()
    te = llava.LlavaForConditionalGeneration.from_synthetic(seed=0, device=dev)
    res["llava_build_s"] = time.time() - t0
    res["llava_mem_gb"] = torch.cuda.memory_allocated() / 2 ** 30
    L, n_img = 359 + 575, 576
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(1000, 100000, (1, L), generator=g)
    ids[:, 5:5 + n_img] = te.config.image_token_index
    mask = torch.ones(1, L, dtype=torch.int64)
    mask[:, 700:] = 0
    ids[:, 700:] = te.config.pad_token_id
    pos = (mask.cumsum(-1) - 1).masked_fill_(mask == 0, 1)
    px = torch.randn(1, 3, 336, 336, generator=g)
    out, ms = timed(lambda: te(input_ids=ids.to(dev), attention_mask=mask.to(dev), position_ids=pos.to(dev), pixel_values=px.to(dev),
                               output_hidden_states=True).hidden_states[-3])
    res["llava"] = dict(tokens=L, ms=ms, finite=bool(torch.isfinite(out[:, :700]).all()), shape=list(out.shape))
    print(res, flush=True)
    del te, out
    torch.cuda.empty_cache()
    clip = encoders.CLIPTextModel.from_synthetic(seed=0, device=dev, **encoders.CLIP_L_TEXT)
    cid = torch.randint(3, 49000, (1, 77), generator=g)
    cid[:, 20:] = 49407
    o, ms = timed(lambda: clip(cid.to(dev)).pooler_output)
    res["clip_l_text"] = dict(ms=ms, finite=bool(torch.isfinite(o).all()), shape=list(o.shape))
    del clip
    torch.cuda.empty_cache()
    vae = AutoencoderKLHunyuanVideo.from_synthetic(seed=0, device=dev)
    x = torch.rand(1, 3, 1, 720, 1280, generator=torch.Generator(device=dev).manual_seed(2), device=dev) * 2 - 1
    torch.cuda.reset_peak_memory_stats()
    m, ms = timed(lambda: vae.encode(x).latent_dist.mode(), n=1)
    res["hunyuan_vae_encode_720p_frame"] = dict(ms=ms, finite=bool(torch.isfinite(m).all()), shape=list(m.shape),
                                                peak_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
    z = torch.randn(1, 16, 3, 30, 40, generator=torch.Generator(device=dev).manual_seed(3), device=dev)
    v, ms = timed(lambda: vae.decode(z).sample, n=1)
    res["hunyuan_vae_decode_9x240x320"] = dict(ms=ms, finite=bool(torch.isfinite(v).all()), shape=list(v.shape))
    print(res, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/r02_hunyuan_aux_fullsize.json", "w"), indent=1)


if __name__ == "__main__":
    main()
