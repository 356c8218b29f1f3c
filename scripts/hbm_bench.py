"""Achieved HBM bandwidth of the HBM-bound hot-path kernels, at the config sizes (launch-latency-bound) and at saturating
synthetic sizes: down_up, gaussian_blur, fused CFG + scheduler step, LayerNorm, RMSNorm + RoPE.
Algorithmic bytes = one read + one write of the tensor (+ the state tensors of the scheduler step).  CUDA events, 20 reps.

    python scripts/hbm_bench.py            # prints one line per kernel; peak from MEASURED_PEAKS.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lp_utils
from alg_b200 import ops, schedulers as S


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


peak = 6500.0
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p)).get("hbm_gbs", peak)


def report(name, nbytes, ms):
    gbs = nbytes / ms / 1e6
    print(f"{name:58s} {nbytes / 1e6:9.1f} MB {ms * 1e3:9.1f} us {gbs:8.0f} GB/s  {gbs / peak:5.2f} of {peak:.0f}", flush=True)


dev = "cuda"
ONLY = os.environ.get("HBM_ONLY", "")  # "norms": only the DiT norm lines


def filters_and_scheduler():
    # ---- down_up (lp_utils.py:49-54): Wan config shape, and 8192 planes
    for planes, tag in ((420, "Wan config [1,20,21,60,104] fp32"), (8192, "8192 planes 60x104 fp32")):
        x = torch.randn(1, planes, 1, 60, 104, device=dev)
        report(f"down_up f=0.4 {tag}", 2 * x.numel() * 4, timed(lambda: lp_utils.apply_low_pass_filter(x, "down_up", 0.0, 0.0, 0.4)))
    x = torch.randn(1, 4096, 1, 90, 160, device=dev).bfloat16()
    report("down_up f=0.625 4096 planes 90x160 bf16", 2 * x.numel() * 2, timed(lambda: lp_utils.apply_low_pass_filter(x, "down_up", 0.0, 0.0, 0.625)))
    # ---- gaussian_blur (lp_utils.py:40-47): Cog config shape, and 768 planes
    for planes, tag in ((3, "Cog config [1,3,480,720] bf16"), (768, "768 planes 480x720 bf16")):
        x = torch.randn(1, planes, 480, 720, device=dev).bfloat16()
        report(f"gaussian k=13 sigma=15 {tag}", 2 * x.numel() * 2, timed(lambda: lp_utils.apply_low_pass_filter(x, "gaussian_blur", 15.0, 0.02734375, 0.25)))
    x = torch.randn(1, 768, 480, 832, device=dev)  # fp32 (Wan's pixel-space ALG filters the fp32 image): the separable kernel
    report("gaussian k=13 sigma=15 768 planes 480x832 fp32 (separable)", 2 * x.numel() * 4, timed(lambda: lp_utils.apply_low_pass_filter(x, "gaussian_blur", 15.0, 0.02734375, 0.25)))
    del x
    # ---- fused CFG + UniPC step (wan:919-927): E = 2 096 640 (config) and 64x that
    for mult, tag in ((1, "Wan config E=2.1M"), (64, "E=134M")):
        E = 2096640 * mult
        s = S.UniPCMultistepScheduler(flow_shift=5.0)
        s.set_timesteps(50, device=dev)
        x = torch.randn(E, device=dev)
        noise = torch.randn(3, E, device=dev).bfloat16()
        for _ in range(3):
            s.step_cfg(noise, 5.0, x)  # reach the order-2 + corrector steady state

        def step():
            s._step_index = 10
            s.step_cfg(noise, 5.0, x)

        report(f"CFG + UniPC step 3-pass order-2 {tag}", (3 * 2 + 4 * 4 + 3 * 4) * E, timed(step))
        del x, noise, s


if ONLY != "norms":
    filters_and_scheduler()
# ---- DiT norms at the Wan 3-pass shape
M, d = 98280, 5120
x = torch.randn(M, d, device=dev).bfloat16()
sc, sh = torch.randn(d, device=dev), torch.randn(d, device=dev)
out = torch.empty_like(x)
report("layer_norm + modulate 98280x5120 bf16", 2 * x.numel() * 2, timed(lambda: ops.layer_norm(x, eps=1e-6, scale=sc, shift=sh, out=out)))
w = torch.randn(128, device=dev).bfloat16()
ang = torch.rand(32760, 64, device=dev) * 6.28
cos, sin = ang.cos().repeat_interleave(2, dim=1).contiguous(), ang.sin().repeat_interleave(2, dim=1).contiguous()
report("head RMSNorm + RoPE 98280 x (40 x 128) bf16", 2 * x.numel() * 2,
       timed(lambda: ops.head_norm_rope(x, 40, 128, norm_kind=1, weight=w, cos=cos, sin=sin, rows_per_batch=32760)))
# the Wan engine's RMSNorm across heads (+ RoPE from the fp64 axis tables), in place (alg_wan_rms_norm_rope)
import ctypes as C
from alg_b200 import _lib
L = _lib.lib()
wd = torch.randn(d, device=dev).bfloat16()
tabs = [torch.stack([a.cos(), a.sin()], dim=-1).contiguous() for a in
        (torch.rand(21, 22, device=dev, dtype=torch.float64) * 6.28, torch.rand(30, 21, device=dev, dtype=torch.float64) * 6.28,
         torch.rand(52, 21, device=dev, dtype=torch.float64) * 6.28)]
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for rope in (True, False):
    ptrs = [C.c_void_p(t.data_ptr()) for t in tabs] if rope else [None, None, None]
    fn = lambda: L.alg_wan_rms_norm_rope(C.c_void_p(x.data_ptr()), M, d, 128, C.c_float(1e-6), C.c_void_p(wd.data_ptr()), *ptrs,
                                         22, 21, 21, 21, 30, 52, st)
    report(f"Wan RMSNorm across heads{' + RoPE' if rope else ''} 98280x5120 bf16", 2 * x.numel() * 2, timed(fn))
del x, out
if ONLY == "norms":
    sys.exit(0)
# ---- CogVideoX VAE encoder blocks at its 128-channel level (480x720 frame)
H, W, Cc = 480, 720, 128
a = torch.randn(H * W, Cc, device=dev).bfloat16()
gw, gb = torch.randn(Cc, device=dev).bfloat16(), torch.randn(Cc, device=dev).bfloat16()
go = torch.empty_like(a)
# GroupNorm reads the activation twice (statistics, then apply) and writes it once
report("group_norm(32) + SiLU 345600x128 bf16 (2 reads + 1 write)", 3 * a.numel() * 2,
       timed(lambda: ops.group_norm(a, 32, gw, gb, eps=1e-6, silu=True, out=go)))
ws = torch.empty(H * W * 9 * Cc, device=dev, dtype=torch.bfloat16)
report("im2col 3x3 patches 345600 x (9 x 128) bf16 (1 read + 9 writes)", 10 * a.numel() * 2,
       timed(lambda: ops.im2col(a, 1, H, W, kernel=(1, 3, 3), pad_top=1, pad_left=1, out=ws)))
