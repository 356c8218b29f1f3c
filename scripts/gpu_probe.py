"""First-contact GPU diagnostics: each section runs in its own subprocess with a timeout so that a trap or hang in
one kernel cannot take the others (or the box) down.  Usage: python scripts/gpu_probe.py [section ...]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sec_env():
    import torch
    print("torch", torch.__version__, "cuda", torch.version.cuda, "gpus", torch.cuda.device_count())
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    print("cpu_count", os.cpu_count(), "ref exists", os.path.exists("/root/reference"))
    os.system("nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv")
    os.system("nproc; free -g | head -2")


def _rel(a, b):
    import torch
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def sec_lowpass():
    import numpy as np
    import torch
    from alg_b200 import lowpass
    g = np.load(os.path.join(ROOT, "tests/golden/lp_down_up.npz"))
    for n in sorted({k[:-2] for k in g.files if k.endswith("_x")}):
        x = torch.from_numpy(g[n + "_x"]).cuda()
        y = lowpass.apply_low_pass_filter(x, "down_up", 0.0, 0.0, float(g[n + "_f"]))
        print("down_up", n, tuple(x.shape), "rel", _rel(y.cpu(), torch.from_numpy(g[n + "_y"])))
    g = np.load(os.path.join(ROOT, "tests/golden/lp_gaussian.npz"))
    for n in sorted({k[:-2] for k in g.files if k.endswith("_x")}):
        dt = getattr(torch, str(g[n + "_dtype"]))
        k = g[n + "_k"]
        k = float(k) if k.dtype == np.float64 else int(k)
        x = torch.from_numpy(g[n + "_x"]).to(dt).cuda()
        y = lowpass.apply_low_pass_filter(x, "gaussian_blur", float(g[n + "_sigma"]), k, 0.0)
        ref = torch.from_numpy(g[n + "_y"])
        d = (y.float().cpu() - ref)
        print("gauss", n, dt, "rel", _rel(y.cpu(), ref), "frac_ne", float((d != 0).float().mean()), "max", float(d.abs().max()))
    # large plane -> generic path
    x = torch.randn(1, 3, 480, 832, device="cuda")
    y = lowpass.apply_low_pass_filter(x, "down_up", 0.0, 0.0, 0.25)
    import torch.nn.functional as F
    r = F.interpolate(F.interpolate(x, size=(120, 208), mode="bilinear", antialias=True), size=(480, 832), mode="bilinear", antialias=True)
    print("down_up generic rel vs torch", _rel(y, r))
    xb = torch.randn(1, 16, 13, 60, 90, device="cuda").bfloat16()
    yb = lowpass.apply_low_pass_filter(xb, "down_up", 0.0, 0.0, 0.25)
    v = xb.view(13 * 1, 16, 60, 90)
    rb = F.interpolate(F.interpolate(v, size=(15, 22), mode="bilinear", antialias=True), size=(60, 90), mode="bilinear", antialias=True).view_as(xb)
    d = (yb.float() - rb.float())
    print("down_up bf16 vs torch cuda: rel", _rel(yb, rb), "frac_ne", float((d != 0).float().mean()))
    # timing on the Wan shape
    x = torch.randn(1, 20, 21, 60, 104, device="cuda")
    for _ in range(3):
        lowpass.apply_low_pass_filter(x, "down_up", 0.0, 0.0, 0.4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20):
        lowpass.apply_low_pass_filter(x, "down_up", 0.0, 0.0, 0.4)
    e1.record(); torch.cuda.synchronize()
    print("down_up wan shape us/call", e0.elapsed_time(e1) * 1000 / 20)
    x = torch.randn(8192, 60, 104, device="cuda").view(1, 8192, 1, 60, 104)
    for _ in range(3):
        lowpass.apply_low_pass_filter(x, "down_up", 0.0, 0.0, 0.4)
    e0.record()
    for _ in range(10):
        lowpass.apply_low_pass_filter(x, "down_up", 0.0, 0.0, 0.4)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("down_up 8192 planes ms", ms, "GB/s", 2 * x.numel() * 4 / ms / 1e6)


def sec_aa():
    import torch
    import torch.nn.functional as F
    from alg_b200 import lowpass
    print("ALG_AA_VARIANT", os.environ.get("ALG_AA_VARIANT"))
    torch.manual_seed(0)
    for dt in (torch.bfloat16, torch.float16):
        for shape, f in (((1, 16, 13, 60, 90), 0.25), ((1, 4, 2, 60, 104), 0.4), ((2, 3, 31, 45), 0.6), ((1, 2, 64, 64), 0.5)):
            x = torch.randn(shape, device="cuda").to(dt)
            y = lowpass.apply_low_pass_filter(x, "down_up", 0.0, 0.0, f)
            H, W = shape[-2:]
            h1, w1 = max(1, int(round(H * f))), max(1, int(round(W * f)))
            v = x.reshape(-1, 1, H, W)
            small = F.interpolate(v, size=(h1, w1), mode="bilinear", antialias=True)
            r = F.interpolate(small, size=(H, W), mode="bilinear", antialias=True).view(shape)
            d = y.float() - r.float()
            print(dt, shape, f, "frac_ne", float((d != 0).float().mean()), "max", float(d.abs().max()), "rel", _rel(y, r))


def sec_sched():
    import torch
    from alg_b200 import schedulers as S
    from oracle import sched_oracle as O
    torch.manual_seed(0)
    E = (1, 16, 5, 12, 20)
    for n_pass in (3, 2):
        s = S.UniPCMultistepScheduler(flow_shift=5.0); s.set_timesteps(12, device="cuda")
        o = O.UniPCOracle(flow_shift=5.0); o.set_timesteps(12)
        x = torch.randn(E); xg = x.cuda()
        worst = 0
        for i in range(12):
            npred = torch.randn((n_pass,) + E[1:]).bfloat16()
            noise = O.cfg_combine(npred.view((n_pass, 1) + E[1:]).squeeze(1).unsqueeze(1).reshape((n_pass,) + E[1:]), 5.0)
            x = o.step(noise, x)
            xg = s.step_cfg(npred.cuda(), 5.0, xg)
            diff = (xg.cpu() - x).abs().max().item()
            worst = max(worst, diff)
        print("unipc n_pass", n_pass, "max abs diff over 12 steps", worst, "bit-exact", worst == 0.0)
    d = S.CogVideoXDDIMScheduler(); d.set_timesteps(10, device="cuda")
    od = O.CogDDIMOracle(); od.set_timesteps(10)
    x = torch.randn(E).bfloat16(); xg = x.cuda()
    worst = 0
    for t in od.timesteps:
        npred = torch.randn((3,) + E[1:]).bfloat16()
        noise = O.cfg_combine(npred, 6.0, fp32=True)
        x = od.step(noise, int(t), x).to(torch.bfloat16)
        xg = d.step_cfg(npred.cuda(), 6.0, int(t), xg)
        worst = max(worst, (xg.float().cpu() - x.float()).abs().max().item())
    print("ddim max abs diff", worst)
    e = S.FlowMatchEulerDiscreteScheduler(shift=7.0); e.set_timesteps(device="cuda", sigmas=torch.linspace(1, 0, 9)[:-1].numpy())
    oe = O.FlowEulerOracle(shift=7.0); oe.set_timesteps(8, sigmas=torch.linspace(1, 0, 9)[:-1].numpy())
    x = torch.randn(E); first = torch.randn(1, 16, 1, 12, 20); xg = x.cuda()
    worst = 0
    for i in range(8):
        npred = torch.randn((2,) + E[1:]).bfloat16()
        noise = O.cfg_combine(npred, 6.0)
        st = oe.step(noise[:, :, 1:], x[:, :, 1:])
        x = torch.cat([first, st], dim=2)
        xg = e.step_cfg_frames(npred.cuda(), 6.0, xg, first.cuda())
        worst = max(worst, (xg.cpu() - x).abs().max().item())
    print("euler max abs diff", worst)


def sec_gemm():
    import torch
    from alg_b200 import ops, _lib
    torch.manual_seed(0)
    def check(M, N, K, epi=0, bias=True, per_row=False, f32=False):
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        b = None
        if bias:
            b = torch.randn(M if per_row else N, device="cuda").bfloat16()
        res = gate = None
        if epi in (2, 3):
            res = torch.randn(M, N, device="cuda").bfloat16()
        if epi == 2:
            gate = torch.randn(1, N, device="cuda")
        out = ops.gemm(a, w, b, epilogue=epi, residual=res, gate=gate, bias_per_row=per_row,
                       out_dtype=torch.float32 if f32 else torch.bfloat16)
        torch.cuda.synchronize()
        ref = a.float() @ w.float().t()
        if bias:
            ref = ref + (b.float()[:, None] if per_row else b.float()[None, :])
        if epi == 1:
            ref = torch.nn.functional.gelu(ref.bfloat16().float(), approximate="tanh")
        elif epi == 4:
            ref = torch.nn.functional.gelu(ref.bfloat16().float())
        elif epi == 2:
            ref = res.float() + ref.bfloat16().float() * gate
        elif epi == 3:
            ref = res.float() + ref.bfloat16().float()
        err = _rel(out, ref)
        print(f"gemm M{M} N{N} K{K} epi{epi} bias{bias} per_row{per_row} f32{f32}: rel {err:.3e}", "OK" if err < 6e-3 else "FAIL")
        return err
    check(128, 256, 64, bias=False, f32=True)
    check(128, 256, 64)
    check(256, 512, 128)
    check(128, 256, 256, bias=False, f32=True)
    check(300, 256, 512)
    check(128, 128, 128)
    check(128, 64, 128)
    check(1000, 1280, 1280, epi=4)
    check(257, 5120, 1280)
    check(1, 30720, 5120)
    check(512, 5120, 4096, epi=1)
    check(777, 5120, 144)
    check(640, 72, 512)
    check(5120, 3000, 5120, per_row=True)
    check(5120, 257, 5120, per_row=True)
    check(4096, 5120, 5120, epi=2)
    check(4096, 5120, 5120, epi=3)
    check(4096, 64, 5120)
    # timing at Wan shapes
    for (M, N, K, epi) in ((65520, 5120, 5120, 0), (65520, 13824, 5120, 1), (65520, 5120, 13824, 2)):
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
        b = torch.randn(N, device="cuda").bfloat16()
        res = torch.randn(M, N, device="cuda").bfloat16() if epi == 2 else None
        gate = torch.randn(1, N, device="cuda") if epi == 2 else None
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for _ in range(2):
            ops.gemm(a, w, b, epilogue=epi, residual=res, gate=gate, out=out)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5):
            ops.gemm(a, w, b, epilogue=epi, residual=res, gate=gate, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        for _ in range(2):
            torch.nn.functional.linear(a, w, b)
        e0.record()
        for _ in range(5):
            torch.nn.functional.linear(a, w, b)
        e1.record(); torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 5
        print(f"gemm time M{M} N{N} K{K} epi{epi}: {ms:.3f} ms {2*M*N*K/ms/1e9:.1f} TFLOP/s | cuBLAS {ms2:.3f} ms {2*M*N*K/ms2/1e9:.1f} TFLOP/s")


def sec_attn():
    import torch
    import torch.nn.functional as F
    from alg_b200 import ops
    torch.manual_seed(0)
    def check(B, H, D, Nq, Nkv, accumulate=False, scale_q=1.0):
        q = (torch.randn(B, Nq, H, D, device="cuda") * scale_q).bfloat16()
        k = torch.randn(B, Nkv, H, D, device="cuda").bfloat16()
        v = torch.randn(B, Nkv, H, D, device="cuda").bfloat16()
        pad = (Nkv + 7) // 8 * 8
        vt = torch.zeros(B, H, D, pad, device="cuda", dtype=torch.bfloat16)
        vt[..., :Nkv] = v.permute(0, 2, 3, 1)
        out = None
        if accumulate:
            base = torch.randn(B, Nq, H, D, device="cuda").bfloat16()
            out = base.clone()
        o = ops.attention(q, k, vt, n_kv=Nkv, out=out, accumulate=accumulate)
        torch.cuda.synchronize()
        ref = F.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)).transpose(1, 2)
        if accumulate:
            ref = ref.bfloat16().float() + base.float()
        err = _rel(o, ref)
        print(f"attn B{B} H{H} D{D} Nq{Nq} Nkv{Nkv} acc{accumulate} sq{scale_q}: rel {err:.3e}", "OK" if err < 1e-2 else "FAIL")
    check(1, 1, 128, 256, 128)
    check(1, 1, 128, 256, 256)
    check(1, 2, 128, 256, 512)
    check(2, 3, 128, 512, 1024)
    check(1, 2, 128, 300, 257)
    check(1, 2, 128, 100, 77)
    check(1, 2, 128, 1000, 2000, scale_q=4.0)
    check(1, 2, 128, 520, 512, accumulate=True)
    check(1, 2, 64, 256, 256)
    check(2, 3, 64, 700, 1000)
    check(1, 4, 128, 4096, 4096, scale_q=3.0)
    # timing at the Wan self-attention shape (one pass)
    B, H, D, N = 1, 40, 128, 32760
    q = torch.randn(B, N, H, D, device="cuda").bfloat16()
    k = torch.randn(B, N, H, D, device="cuda").bfloat16()
    vt = torch.randn(B, H, D, N, device="cuda").bfloat16()
    o = torch.empty_like(q)
    for _ in range(2):
        ops.attention(q, k, vt, out=o)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3):
        ops.attention(q, k, vt, out=o)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    fl = 4 * B * H * N * N * D
    print(f"attention wan shape: {ms:.2f} ms {fl/ms/1e9:.1f} TFLOP/s")
    qt, kt, v = q.transpose(1, 2), k.transpose(1, 2), vt.transpose(2, 3)
    from torch.nn.attention import SDPBackend, sdpa_kernel
    for be in (SDPBackend.FLASH_ATTENTION, SDPBackend.CUDNN_ATTENTION, SDPBackend.EFFICIENT_ATTENTION):
        try:
            with sdpa_kernel(be):
                for _ in range(2):
                    F.scaled_dot_product_attention(qt, kt, v)
                e0.record()
                for _ in range(3):
                    F.scaled_dot_product_attention(qt, kt, v)
                e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f"torch SDPA {be}: {ms:.2f} ms {fl/ms/1e9:.1f} TFLOP/s")
        except Exception as ex:
            print("torch SDPA", be, "failed:", str(ex)[:100])


def sec_wan_tiny():
    import torch
    from alg_b200 import wan
    from oracle import wan_oracle as W
    torch.manual_seed(0)
    cfgo = W.tiny_config(num_attention_heads=2, num_layers=2)
    cfg = dict(num_attention_heads=2, attention_head_dim=128, text_dim=64, freq_dim=256, ffn_dim=512, num_layers=2,
               image_dim=64, text_len=32)
    m = wan.WanTransformer3DModel.from_synthetic(seed=3, device="cuda", **cfg)
    sd = m.state_dict()
    T, H, Wd = 3, 16, 24
    N = T * (H // 2) * (Wd // 2)
    d = 256
    lat = torch.randn(16, T, H, Wd, device="cuda")
    c0 = torch.randn(20, T, H, Wd, device="cuda"); c1 = torch.randn(20, T, H, Wd, device="cuda")
    neg = torch.randn(32, 64, device="cuda").bfloat16(); pos = torch.randn(32, 64, device="cuda").bfloat16()
    img = torch.randn(9, 64, device="cuda").bfloat16()
    dbg = m.enable_debug(3 * N * d * 2 * 4 + 7 * d * 2 + 1024)
    out = m.forward_passes([lat] * 3, [c0, c1, c1], [neg, neg, pos], img, 987)
    torch.cuda.synchronize()
    x = torch.cat([torch.stack([lat] * 3), torch.stack([c0, c1, c1])], dim=1).bfloat16()
    t = torch.tensor([987] * 3, device="cuda")
    text = torch.stack([neg, neg, pos]); im = img[None].repeat(3, 1, 1)
    ref, inter = W.forward(sd, cfgo, x, t, text, im, return_intermediates=True)
    sd32 = {k: v.float() for k, v in sd.items()}
    ref32, inter32 = W.forward(sd32, cfgo, x.float(), t, text.float(), im.float(), return_intermediates=True)
    dbgb = dbg.view(torch.bfloat16)
    off = 0
    def take(n):
        nonlocal off
        r = dbgb[off:off + n]; off += n
        return r
    patch = take(3 * N * d).view(3, N, d)
    temb = take(d); tproj = take(6 * d)
    print("patch rel vs bf16 oracle", _rel(patch, inter["patch"]), "vs fp32", _rel(patch, inter32["patch"]))
    print("temb rel", _rel(temb, inter["temb"][0]), "tproj rel", _rel(tproj, inter["tproj"][0].reshape(-1)))
    for i in range(2):
        blk = take(3 * N * d).view(3, N, d)
        print(f"block{i} rel vs bf16 oracle", _rel(blk, inter[f"block{i}"]), "vs fp32", _rel(blk, inter32[f"block{i}"]),
              "| bf16 oracle vs fp32", _rel(inter[f"block{i}"], inter32[f"block{i}"]))
    print("out rel vs bf16 oracle", _rel(out, ref), "vs fp32 oracle", _rel(out, ref32), "| bf16 oracle vs fp32", _rel(ref, ref32))


SECTIONS = {k[4:]: v for k, v in list(globals().items()) if k.startswith("sec_")}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        SECTIONS[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(SECTIONS)
    for n in names:
        print(f"===== {n} =====", flush=True)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", n], timeout=300, cwd=ROOT)
            print(f"----- {n}: exit {r.returncode} in {time.time()-t0:.1f}s", flush=True)
        except subprocess.TimeoutExpired:
            print(f"----- {n}: TIMEOUT", flush=True)
