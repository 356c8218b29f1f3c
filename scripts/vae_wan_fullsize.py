"""Wan2.1 VAE at the BASELINE configs[1] size on one B200: native decode of [1, 16, 21, 60, 104] -> 81 x 480 x 832 and native
encode of the 81-frame condition clip (wan:429-434), timed with CUDA events, checked against the eager oracle
(oracle/wan_vae_oracle.py, fp32 with TF32 off and, as PyTorch's default, TF32 on) on a 2-latent-frame slice at full resolution.
Writes gpurun_out/r02_vae_wan_fullsize.json."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alg_b200.vae_wan import AutoencoderKLWan  # noqa: E402
from oracle import wan_vae_oracle as V  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def timed(fn, n=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / n


def main():
    dev = "cuda"
    cfg = dict(V.WAN21_VAE)
    sd = V.make_weights(cfg, seed=0, device=dev)
    vae = AutoencoderKLWan(**cfg).load_state_dict(sd)
    res = {"config": "Wan2.1 VAE, 81 x 480 x 832 (latent 21 x 60 x 104), float32 via bf16 3-term split"}
    g = torch.Generator(device=dev).manual_seed(1)
    # parity slice at full resolution: 2 latent frames -> 5 frames
    z = torch.randn(1, 16, 2, 60, 104, generator=g, device=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref, t_ref = timed(lambda: V.decode(z, sd, cfg, torch.float32))
    out, t_nat = timed(lambda: vae.decode(z).sample)
    torch.backends.cudnn.allow_tf32 = True
    ref_tf32, t_ref_tf32 = timed(lambda: V.decode(z, sd, cfg, torch.float32))
    res["decode_slice"] = dict(latent_frames=2, native_vs_eager_fp32=rel(out, ref), eager_tf32_vs_eager_fp32=rel(ref_tf32, ref),
                               native_ms=t_nat, eager_fp32_ms=t_ref, eager_tf32_ms=t_ref_tf32)
    print(res["decode_slice"], flush=True)
    x = torch.rand(1, 3, 5, 480, 832, generator=g, device=dev) * 2 - 1
    torch.backends.cudnn.allow_tf32 = False
    mref, t_ref = timed(lambda: V.encode_moments(x, sd, cfg, torch.float32))
    mout, t_nat = timed(lambda: vae.encode(x).latent_dist.parameters)
    torch.backends.cudnn.allow_tf32 = True
    mref_tf32, t_ref_tf32 = timed(lambda: V.encode_moments(x, sd, cfg, torch.float32))
    res["encode_slice"] = dict(frames=5, native_vs_eager_fp32=rel(mout, mref), eager_tf32_vs_eager_fp32=rel(mref_tf32, mref),
                               native_ms=t_nat, eager_fp32_ms=t_ref, eager_tf32_ms=t_ref_tf32)
    print(res["encode_slice"], flush=True)
    del ref, out, ref_tf32, mref, mout, mref_tf32
    torch.cuda.empty_cache()
    # full clip
    z = torch.randn(1, 16, 21, 60, 104, generator=g, device=dev)
    torch.cuda.reset_peak_memory_stats()
    video, t_dec = timed(lambda: vae.decode(z).sample)
    res["decode_full"] = dict(ms=t_dec, frames=int(video.shape[2]), finite=bool(torch.isfinite(video).all()),
                              peak_gb=torch.cuda.max_memory_allocated() / 2 ** 30, frames_per_s=video.shape[2] / (t_dec / 1e3))
    print(res["decode_full"], flush=True)
    if os.environ.get("VAE_EAGER_FULL", "1") == "1":
        torch.backends.cudnn.allow_tf32 = True
        _, t_eager = timed(lambda: V.decode(z, sd, cfg, torch.float32))
        res["decode_full"]["eager_tf32_ms"] = t_eager
        print("eager tf32 decode", t_eager, flush=True)
    del video
    torch.cuda.empty_cache()
    x = torch.zeros(1, 3, 81, 480, 832, device=dev)
    x[:, :, 0] = torch.rand(1, 3, 480, 832, generator=g, device=dev) * 2 - 1
    torch.cuda.reset_peak_memory_stats()
    m, t_enc = timed(lambda: vae.encode(x).latent_dist.mode())
    res["encode_full"] = dict(ms=t_enc, latent=list(m.shape), finite=bool(torch.isfinite(m).all()),
                              peak_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
    print(res["encode_full"], flush=True)
    if os.environ.get("VAE_EAGER_FULL", "1") == "1":
        _, t_eager = timed(lambda: V.encode_moments(x, sd, cfg, torch.float32))
        res["encode_full"]["eager_tf32_ms"] = t_eager
        print("eager tf32 encode", t_eager, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/r02_vae_wan_fullsize.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
