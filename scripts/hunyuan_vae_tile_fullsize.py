"""One temporal tile of AutoencoderKLHunyuanVideo.decode at the BASELINE configs[3] size (5 latent frames of 90 x 160 -> 17 frames of
720 x 1280; the 129-frame clip is 11 such tiles, hy:1292): time, peak memory, finiteness."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alg_b200.vae_hunyuan import AutoencoderKLHunyuanVideo  # noqa: E402

vae = AutoencoderKLHunyuanVideo.from_synthetic(seed=0, device="cuda")
z = torch.randn(1, 16, 5, 90, 160, device="cuda")
torch.cuda.synchronize()
t0 = time.time()
v = vae._decode_tile(z[0])
torch.cuda.synchronize()
print("tile decode s", round(time.time() - t0, 2), tuple(v.shape), bool(torch.isfinite(v).all()), "peak GB",
      round(torch.cuda.max_memory_allocated() / 2 ** 30, 1), flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "full":  # the whole BASELINE configs[3] clip: 33 latent frames -> 129 frames of 720 x 1280
    import json
    z = torch.randn(1, 16, 33, 90, 160, device="cuda")
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    t0 = time.time()
    v = vae.decode(z).sample
    torch.cuda.synchronize()
    dt = time.time() - t0
    res = dict(clip="33 x 90 x 160 latents -> 129 x 720 x 1280 (hy:1292), 11 temporal tiles of 5 latent frames, cross-faded", seconds=dt,
               shape=list(v.shape), finite=bool(torch.isfinite(v).all()), peak_gb=torch.cuda.max_memory_allocated() / 2 ** 30,
               frames_per_s=v.shape[2] / dt)
    print(res, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/r02_hunyuan_vae_decode_fullsize.json", "w"), indent=1)
