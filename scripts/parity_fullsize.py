"""Full-size parity + the eager-PyTorch GPU bar (VERDICT r1 "next" #1; SURVEY 8(c) protocol, 2.3 row 4).

    python scripts/parity_fullsize.py wan cog hunyuan --out gpurun_out/r02_parity_fullsize.json

For each model at its BASELINE.json config size and FULL depth:

  * teacher-forced denoise steps (Wan: one 3-pass ALG step at schedule index 0 and one 2-pass step at index 10; Cog: a
    3-pass and a 2-pass step; Hunyuan: the single-pass ALG step) run through the native engine AND through the eager
    bf16 oracle (``oracle/*_oracle.forward`` = PyTorch ops: cuBLAS linears, SDPA) on the same tensors; both sides then
    apply CFG + the scheduler update from identical, synced scheduler history.  Recorded: relative L2 of ``noise_pred``
    and of the stepped latents (the north_star contract: <= 1e-3), and the eager step time (the GPU bar);
  * the fp32 ground truth at a reduced depth, full token count: err(engine vs fp32) against err(eager bf16 vs fp32);
  * the per-op bars at the config shapes: torch SDPA (flash / cuDNN / efficient) and cuBLAS ``F.linear``.

Checker only: everything under ``oracle/`` is test infrastructure, nothing here is on the product path.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


class Timer:
    def __enter__(self):
        torch.cuda.synchronize()
        self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.e0.record()
        return self

    def __exit__(self, *a):
        self.e1.record()
        torch.cuda.synchronize()
        self.ms = self.e0.elapsed_time(self.e1)


def chunked_fp32_attention(q, k, v, rows=4096):
    """softmax(q k^T / sqrt(d)) v in plain fp32 matmuls (no TF32, no fused kernel): the ground truth for N ~ 1e5 tokens."""
    assert q.dtype == torch.float32
    B, H, N, D = q.shape
    out = torch.empty_like(q)
    scale = D ** -0.5
    for b in range(B):
        for h in range(H):
            kt = k[b, h].t().contiguous()
            for r0 in range(0, N, rows):
                s = (q[b, h, r0:r0 + rows] @ kt) * scale
                out[b, h, r0:r0 + rows] = torch.softmax(s, dim=-1) @ v[b, h]
    return out


def sdpa_backend(name):
    from torch.nn.attention import SDPBackend, sdpa_kernel
    return sdpa_kernel({"flash": SDPBackend.FLASH_ATTENTION, "cudnn": SDPBackend.CUDNN_ATTENTION,
                        "efficient": SDPBackend.EFFICIENT_ATTENTION, "math": SDPBackend.MATH}[name])


def time_call(fn, reps=1, warm=1):
    for _ in range(warm):
        fn()
    with Timer() as t:
        for _ in range(reps):
            fn()
    return t.ms / reps


# ======================================================================================================
def op_bars(shapes_attn, shapes_gemm, reps=3):
    """torch SDPA and cuBLAS at the config's shapes, and this repo's kernels beside them."""
    from alg_b200 import ops
    out = {"sdpa": [], "gemm": []}
    g = torch.Generator(device="cuda").manual_seed(1)
    for (B, H, D, N) in shapes_attn:
        q = torch.randn(B, N, H, D, generator=g, device="cuda").bfloat16()
        k = torch.randn(B, N, H, D, generator=g, device="cuda").bfloat16()
        v = torch.randn(B, N, H, D, generator=g, device="cuda").bfloat16()
        npad = (N + 7) // 8 * 8
        vt = torch.zeros(B, H, D, npad, device="cuda", dtype=torch.bfloat16)
        vt[..., :N] = v.permute(0, 2, 3, 1)
        o = torch.empty_like(q)
        fl = 4 * B * H * N * N * D
        row = {"shape": {"B": B, "heads": H, "head_dim": D, "tokens": N}, "flops": fl}
        ms = time_call(lambda: ops.attention(q, k, vt, n_kv=N, out=o), reps)
        row["alg_b200"] = {"ms": ms, "tflops": fl / ms / 1e9}
        qt, kt, vv = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
        for be in ("flash", "cudnn", "efficient"):
            try:
                with sdpa_backend(be):
                    ms = time_call(lambda: F.scaled_dot_product_attention(qt, kt, vv), reps)
                row["torch_" + be] = {"ms": ms, "tflops": fl / ms / 1e9}
            except Exception as ex:  # backend not available for this shape / build
                row["torch_" + be] = {"error": str(ex).splitlines()[0][:160]}
        out["sdpa"].append(row)
        del q, k, v, vt, o
    for (M, N, K) in shapes_gemm:
        a = torch.randn(M, K, generator=g, device="cuda").bfloat16()
        w = (torch.randn(N, K, generator=g, device="cuda") * 0.02).bfloat16()
        b = torch.randn(N, generator=g, device="cuda").bfloat16()
        o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        fl = 2 * M * N * K
        ms = time_call(lambda: ops.gemm(a, w, b, out=o), reps)
        ms2 = time_call(lambda: F.linear(a, w, b), reps)
        out["gemm"].append({"shape": {"M": M, "N": N, "K": K}, "flops": fl, "alg_b200": {"ms": ms, "tflops": fl / ms / 1e9},
                            "torch_cublas": {"ms": ms2, "tflops": fl / ms2 / 1e9}})
        del a, w, b, o
    return out


# ======================================================================================================
def wan_case(layers=40, fp32_layers=2, eager_backends=("flash", "cudnn"), resolution="480p", log=print):
    import bench
    from alg_b200 import lowpass, wan
    from alg_b200.schedulers import UniPCMultistepScheduler
    from oracle import sched_oracle, wan_oracle as W

    dev = torch.device("cuda")
    if resolution == "720p":
        bench.HEIGHT, bench.WIDTH, bench.H_LAT, bench.W_LAT = 720, 1280, 90, 160
    T, Hl, Wl = bench.T_LAT, bench.H_LAT, bench.W_LAT
    n_tok = T * (Hl // 2) * (Wl // 2)
    res = {"model": f"Wan-I2V-14B {bench.HEIGHT}x{bench.WIDTH}, 81 frames (BASELINE.json configs[1])", "layers": layers,
           "tokens": n_tok, "guidance": bench.GUIDANCE, "flow_shift": bench.FLOW_SHIFT, "steps": []}
    lat0, cond, pos, neg, img = bench.synthetic_inputs(dev, 42)
    # the oracle side's low-pass: the two ATen calls of lp_utils.py:49-54, on the GPU
    h1, w1 = max(1, int(round(Hl * 0.4))), max(1, int(round(Wl * 0.4)))
    c4 = cond.view(-1, 1, Hl, Wl)
    lp_ref = F.interpolate(F.interpolate(c4, size=(h1, w1), mode="bilinear", align_corners=False, antialias=True),
                           size=(Hl, Wl), mode="bilinear", align_corners=False, antialias=True).view_as(cond)
    lp_eng = lowpass.apply_low_pass_filter(cond, "down_up", 0.0, 0.0, 0.4)
    res["lp_rel_l2"] = rel_l2(lp_eng, lp_ref)

    def run_depth(n_layers, fp32):
        t0 = time.time()
        model = wan.WanTransformer3DModel.from_synthetic(seed=0, device=dev, num_layers=n_layers)
        sd = model.state_dict()
        ocfg = W.WanConfig(num_layers=n_layers)
        log(f"[wan] {n_layers}-layer weights in {time.time() - t0:.1f} s")
        sched = UniPCMultistepScheduler(flow_shift=bench.FLOW_SHIFT)
        sched.set_timesteps(bench.STEPS_PER_VIDEO, device=dev)
        osch = sched_oracle.UniPCOracle(flow_shift=bench.FLOW_SHIFT)
        osch.set_timesteps(bench.STEPS_PER_VIDEO)
        ts = sched.timesteps.tolist()
        gfake = torch.Generator(device=dev).manual_seed(7)
        lat = lat0
        out_rows = []
        for idx, n_pass in ((0, 3), (10, 2)):
            # bring both schedulers to `idx` with the SAME fake model outputs (teacher-forced, history synced)
            while sched.step_index < idx:
                fake = torch.randn(lat.shape, generator=gfake, device=dev).bfloat16()
                a = sched.step_cfg(fake, 1.0, lat)
                b = osch.step(fake, lat)
                if not torch.equal(a, b):
                    res["history_sync_max_abs_diff"] = max(res.get("history_sync_max_abs_diff", 0.0), float((a - b).abs().max()))
                lat = b
            conds_e = [cond[0], lp_eng[0], lp_eng[0]] if n_pass == 3 else [cond[0], cond[0]]
            conds_r = [cond, lp_ref, lp_ref] if n_pass == 3 else [cond, cond]
            texts = [neg[0], neg[0], pos[0]] if n_pass == 3 else [neg[0], pos[0]]
            with Timer() as te:
                noise_eng = model.forward_passes([lat[0]] * n_pass, conds_e, texts, img[0], ts[idx])
            with Timer() as te2:  # second call: steady state
                noise_eng = model.forward_passes([lat[0]] * n_pass, conds_e, texts, img[0], ts[idx])
            x = torch.cat([torch.cat([lat] * n_pass), torch.cat(conds_r)], dim=1).bfloat16()
            text = torch.stack(texts)
            tt = torch.tensor([ts[idx]] * n_pass, device=dev)
            im = img.repeat(n_pass, 1, 1)
            row = {"schedule_index": idx, "n_pass": n_pass, "timestep": ts[idx], "engine_ms": te2.ms, "engine_first_call_ms": te.ms}
            noise_ref = None
            with torch.no_grad():
                for be in eager_backends:
                    try:
                        with sdpa_backend(be):
                            if n_layers > 1:  # warm the eager path (cuDNN / cuBLAS plan selection) on the first block only
                                W.forward(sd, W.WanConfig(num_layers=1), x, tt, text, im)
                            with Timer() as tr:
                                nr = W.forward(sd, ocfg, x, tt, text, im)
                        row[f"eager_{be}_ms"] = tr.ms
                        if noise_ref is None:
                            noise_ref, row["eager_reference_backend"] = nr, be
                        else:  # two eager bf16 evaluations of the same model: the floor any bf16 implementation sits on
                            row[f"eager_{be}_vs_reference_rel_l2"] = rel_l2(nr, noise_ref)
                            s_a, s_b = copy.deepcopy(osch), copy.deepcopy(osch)
                            row[f"eager_{be}_vs_reference_latent_rel_l2"] = rel_l2(
                                s_a.step(sched_oracle.cfg_combine(nr, bench.GUIDANCE), lat),
                                s_b.step(sched_oracle.cfg_combine(noise_ref, bench.GUIDANCE), lat))
                        del nr
                    except Exception as ex:
                        row[f"eager_{be}_error"] = str(ex).splitlines()[0][:200]
                    torch.cuda.empty_cache()
            row["noise_finite"] = bool(torch.isfinite(noise_eng).all())
            row["noise_rel_l2"] = rel_l2(noise_eng, noise_ref)
            # CFG + UniPC on both sides, from identical history
            s_e, s_r = copy.copy(sched), copy.deepcopy(osch)
            s_e._state = [t.clone() for t in sched._state] if sched._state is not None else None
            x_eng = s_e.step_cfg(noise_eng, bench.GUIDANCE, lat)
            x_ref = s_r.step(sched_oracle.cfg_combine(noise_ref, bench.GUIDANCE), lat)
            row["latent_rel_l2"] = rel_l2(x_eng, x_ref)
            s_k = copy.copy(sched)
            s_k._state = [t.clone() for t in sched._state] if sched._state is not None else None
            row["sched_kernel_on_oracle_noise_bitexact"] = bool(torch.equal(s_k.step_cfg(noise_ref, bench.GUIDANCE, lat), x_ref))
            row["delta_sigma"] = float(osch.sigmas[idx + 1] - osch.sigmas[idx])
            if fp32:
                sd32 = {k: v.float() for k, v in sd.items()}
                old = W.sdpa
                W.sdpa = lambda q, k, v: chunked_fp32_attention(q, k, v)
                try:
                    with torch.no_grad(), Timer() as t32:
                        n32 = W.forward(sd32, ocfg, x.float(), tt, text.float(), im.float())
                finally:
                    W.sdpa = old
                row["fp32_ms"] = t32.ms
                row["engine_vs_fp32"] = rel_l2(noise_eng, n32)
                row["eager_bf16_vs_fp32"] = rel_l2(noise_ref, n32)
                del sd32, n32
            log(f"[wan L={n_layers}] " + json.dumps(row))
            out_rows.append(row)
            # advance both schedulers with the oracle's noise (teacher)
            comb = sched_oracle.cfg_combine(noise_ref, bench.GUIDANCE)
            a = sched.step_cfg(comb, 1.0, lat)
            b = osch.step(comb, lat)
            lat = b
            del noise_ref, noise_eng, x
            torch.cuda.empty_cache()
        del model, sd
        torch.cuda.empty_cache()
        return out_rows

    if fp32_layers:
        res["fp32_ground_truth"] = {"layers": fp32_layers, "steps": run_depth(fp32_layers, True)}
    res["steps"] = run_depth(layers, False)
    res["max_latent_rel_l2"] = max(r["latent_rel_l2"] for r in res["steps"])
    res["latent_within_1e-3"] = res["max_latent_rel_l2"] <= 1e-3
    return res


# ======================================================================================================
def cog_case(layers=42, fp32_layers=2, eager_backends=("flash", "cudnn"), log=print):
    from alg_b200 import cogvideox
    from alg_b200.schedulers import CogVideoXDDIMScheduler
    from oracle import cog_oracle as Co, sched_oracle

    dev = torch.device("cuda")
    Fr, H, W, steps, gs = 13, 60, 90, 50, 6.0
    res = {"model": "CogVideoX-5b-I2V 480x720, 49 frames (BASELINE.json configs[2])", "layers": layers,
           "tokens": Fr * (H // 2) * (W // 2) + 226, "guidance": gs, "steps": []}
    g = torch.Generator(device=dev).manual_seed(42)
    lat = torch.randn(1, Fr, 16, H, W, generator=g, device=dev).bfloat16()
    img_lat = torch.cat([torch.randn(1, 1, 16, H, W, generator=g, device=dev), torch.zeros(1, Fr - 1, 16, H, W, device=dev)], 1).bfloat16()
    lp = torch.cat([torch.randn(1, 1, 16, H, W, generator=g, device=dev), torch.zeros(1, Fr - 1, 16, H, W, device=dev)], 1).bfloat16()
    pos, neg = (torch.randn(1, 226, 4096, generator=g, device=dev).bfloat16() for _ in range(2))

    def run_depth(n_layers, fp32):
        model = cogvideox.CogVideoXTransformer3DModel.from_synthetic(seed=0, device=dev, num_layers=n_layers)
        sd = model.state_dict()
        ocfg = Co.CogConfig(num_layers=n_layers)
        rope = tuple(r.to(dev) for r in Co.rotary_tables(ocfg, H // 2, W // 2, Fr))
        sched = CogVideoXDDIMScheduler()
        sched.set_timesteps(steps, device=dev)
        osch = sched_oracle.CogDDIMOracle()
        osch.set_timesteps(steps)
        rows = []
        for idx, n_pass in ((0, 3), (10, 2)):
            t = int(sched.timesteps[idx])
            conds = [img_lat, lp, lp] if n_pass == 3 else [img_lat, img_lat]
            x = torch.cat([torch.cat([lat] * n_pass), torch.cat(conds)], dim=2)
            text = torch.cat([neg, neg, pos] if n_pass == 3 else [neg, pos])
            tt = torch.tensor([t] * n_pass, device=dev)
            model(x, text, tt, image_rotary_emb=rope, return_dict=False)
            with Timer() as te:
                noise_eng = model(x, text, tt, image_rotary_emb=rope, return_dict=False)[0]
            row = {"schedule_index": idx, "n_pass": n_pass, "timestep": t, "engine_ms": te.ms}
            noise_ref = None
            with torch.no_grad():
                for be in eager_backends:
                    try:
                        with sdpa_backend(be):
                            Co.forward(sd, ocfg, x, text, tt, rope) if noise_ref is None else None
                            with Timer() as tr:
                                nr = Co.forward(sd, ocfg, x, text, tt, rope)
                        row[f"eager_{be}_ms"] = tr.ms
                        if noise_ref is None:
                            noise_ref, row["eager_reference_backend"] = nr, be
                    except Exception as ex:
                        row[f"eager_{be}_error"] = str(ex).splitlines()[0][:200]
            row["noise_finite"] = bool(torch.isfinite(noise_eng).all())
            row["noise_rel_l2"] = rel_l2(noise_eng, noise_ref)
            x_eng = sched.step_cfg(noise_eng, gs, t, lat)
            x_ref = osch.step(sched_oracle.cfg_combine(noise_ref, gs, fp32=True), t, lat).to(pos.dtype)
            row["latent_rel_l2"] = rel_l2(x_eng, x_ref)
            row["sched_kernel_on_oracle_noise_bitexact"] = bool(torch.equal(sched.step_cfg(noise_ref, gs, t, lat), x_ref))
            if fp32:
                sd32 = {k: v.float() for k, v in sd.items()}
                old = Co.sdpa
                Co.sdpa = lambda q, k, v: chunked_fp32_attention(q, k, v)
                try:
                    with torch.no_grad():
                        n32 = Co.forward(sd32, ocfg, x.float(), text.float(), tt, rope)
                finally:
                    Co.sdpa = old
                row["engine_vs_fp32"] = rel_l2(noise_eng, n32)
                row["eager_bf16_vs_fp32"] = rel_l2(noise_ref, n32)
                del sd32, n32
            log(f"[cog L={n_layers}] " + json.dumps(row))
            rows.append(row)
        del model, sd
        torch.cuda.empty_cache()
        return rows

    if fp32_layers:
        res["fp32_ground_truth"] = {"layers": fp32_layers, "steps": run_depth(fp32_layers, True)}
    res["steps"] = run_depth(layers, False)
    res["max_latent_rel_l2"] = max(r["latent_rel_l2"] for r in res["steps"])
    res["latent_within_2^-8"] = res["max_latent_rel_l2"] <= 2 ** -8  # bf16 latent state: one rounding of the state (cog:1123)
    return res


# ======================================================================================================
def hunyuan_case(layers=20, single_layers=40, fp32_layers=0, eager_backends=("efficient", "cudnn"), log=print):
    from alg_b200 import hunyuan
    from alg_b200.schedulers import FlowMatchEulerDiscreteScheduler
    from oracle import hunyuan_oracle as Ho, sched_oracle
    import numpy as np

    dev = torch.device("cuda")
    T, H, W, steps, L, valid = 33, 90, 160, 30, 256, 180
    res = {"model": "HunyuanVideo-I2V 720x1280, 129 frames, single-pass ALG branch (BASELINE.json configs[3])",
           "layers": [layers, single_layers], "tokens": T * (H // 2) * (W // 2) + valid, "steps": []}
    g = torch.Generator(device=dev).manual_seed(42)
    lat = torch.randn(1, 16, T, H, W, generator=g, device=dev)
    first = torch.randn(1, 16, 1, H, W, generator=g, device=dev)
    text = torch.randn(1, L, 4096, generator=g, device=dev).bfloat16()
    mask = torch.zeros(1, L, device=dev)
    mask[:, :valid] = 1
    pooled = torch.randn(1, 768, generator=g, device=dev).bfloat16()
    model = hunyuan.HunyuanVideoTransformer3DModel.from_synthetic(seed=0, device=dev, num_layers=layers, num_single_layers=single_layers)
    sd = model.state_dict()
    ocfg = Ho.HunyuanConfig(num_layers=layers, num_single_layers=single_layers)
    sched = FlowMatchEulerDiscreteScheduler(shift=7.0)
    sig = np.linspace(1.0, 0.0, steps + 1)[:-1]
    sched.set_timesteps(sigmas=sig, device=dev)
    osch = sched_oracle.FlowEulerOracle(shift=7.0)
    osch.set_timesteps(steps, sigmas=sig)
    idx = 0
    t = sched.timesteps[idx].to(torch.bfloat16).reshape(1)
    gd = torch.tensor([6.0], device=dev).bfloat16() * 1000.0
    x = torch.cat([first, lat[:, :, 1:]], dim=2)  # hy:1232-1235 with s == 0 / identity filter: frame 0 replaced
    model(x, t, text, mask, pooled, gd, return_dict=False)
    with Timer() as te:
        noise_eng = model(x, t, text, mask, pooled, gd, return_dict=False)[0]
    row = {"schedule_index": idx, "n_pass": 1, "engine_ms": te.ms}
    noise_ref = None
    with torch.no_grad():
        for be in eager_backends:
            try:
                with sdpa_backend(be):
                    with Timer() as tr:
                        nr = Ho.forward(sd, ocfg, x.bfloat16(), t, text, mask, pooled, gd)
                row[f"eager_{be}_ms"] = tr.ms
                if noise_ref is None:
                    noise_ref, row["eager_reference_backend"] = nr, be
            except Exception as ex:
                row[f"eager_{be}_error"] = str(ex).splitlines()[0][:200]
            torch.cuda.empty_cache()
    row["noise_finite"] = bool(torch.isfinite(noise_eng).all())
    if noise_ref is not None:
        row["noise_rel_l2"] = rel_l2(noise_eng, noise_ref)
        x_eng = sched.step_cfg_frames(noise_eng, 1.0, torch.cat([first, lat[:, :, 1:]], 2), first)
        stepped = osch.step(noise_ref[:, :, 1:], lat[:, :, 1:])
        x_ref = torch.cat([first, stepped.float()], dim=2)
        row["latent_rel_l2"] = rel_l2(x_eng, x_ref)
    log("[hunyuan] " + json.dumps(row))
    res["steps"].append(row)
    if "latent_rel_l2" in row:
        res["max_latent_rel_l2"] = row["latent_rel_l2"]
        res["latent_within_2^-8"] = row["latent_rel_l2"] <= 2 ** -8
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("models", nargs="*", default=["wan"])
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_parity_fullsize.json"))
    ap.add_argument("--layers", type=int, default=None, help="override the depth (debug)")
    ap.add_argument("--no-bars", action="store_true")
    args = ap.parse_args()
    from alg_b200 import _lib
    _lib.check(_lib.lib().alg_check_device())
    result = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "cases": {}}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)

    def save():
        with open(args.out, "w") as f:
            json.dump(result, f, indent=1)

    log = lambda s: print(s, file=sys.stderr, flush=True)  # noqa: E731
    for m in args.models:
        try:
            with torch.no_grad():
                if m == "wan":
                    result["cases"]["wan"] = wan_case(layers=args.layers or 40, log=log)
                elif m == "wan720":
                    result["cases"]["wan720"] = wan_case(layers=args.layers or 40, fp32_layers=0, resolution="720p", log=log)
                elif m == "cog":
                    result["cases"]["cog"] = cog_case(layers=args.layers or 42, log=log)
                elif m == "hunyuan":
                    result["cases"]["hunyuan"] = hunyuan_case(log=log) if not args.layers else hunyuan_case(args.layers, args.layers, log=log)
        except Exception as ex:
            import traceback
            traceback.print_exc()
            result["cases"][m] = {"error": f"{type(ex).__name__}: {ex}"[:400]}
        torch.cuda.empty_cache()
        save()
    if not args.no_bars:
        try:
            result["op_bars_wan"] = op_bars([(1, 40, 128, 32760)], [(65520, 5120, 5120), (65520, 13824, 5120), (65520, 5120, 13824)])
            result["op_bars_cog"] = op_bars([(2, 48, 64, 17776)], [(35552, 3072, 3072), (35552, 12288, 3072)])
        except Exception as ex:
            result["op_bars_error"] = str(ex)[:300]
        save()
    print(json.dumps(result))


if __name__ == "__main__":
    main()
