"""CPU-side checks: the C-ABI library loads and exports every declared symbol; host mirrors match the reference
fixtures and the oracle; product code never imports the oracle."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from alg_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "alg_b200.h")).read()
    declared = set(re.findall(r"\b(alg_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(_lib.SIGNATURES)
    assert _lib.lib().alg_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of every ABI struct as gcc sees include/alg_b200.h == the ctypes mirror in _lib.py."""
    import subprocess

    from alg_b200 import _lib

    pairs = {"alg_unipc_step_t": _lib.UniPCStep, "alg_dpm_step_t": _lib.DpmStep, "alg_gemm_t": _lib.Gemm, "alg_attention_t": _lib.Attention,
             "alg_wan_config_t": _lib.WanConfig, "alg_layer_norm_t": _lib.LayerNorm,
             "alg_head_norm_rope_t": _lib.HeadNormRope, "alg_patch_src_t": _lib.PatchSrc,
             "alg_im2col_t": _lib.Im2col, "alg_group_norm_t": _lib.GroupNorm}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "alg_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0; }")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in pairs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_gaussian_taps_are_dtype_faithful():
    from alg_b200 import _lib
    from oracle import lp_oracle as O

    L = _lib.lib()
    for k, sigma in ((13, 15.0), (15, 7.5), (5, 1.25), (1, 3.0), (9, 0.6)):
        for code, name in ((0, "float32"), (1, "bfloat16"), (2, "float16")):
            buf = (ctypes.c_float * 64)()
            assert L.alg_gaussian_kernel1d(k, sigma, code, buf) == 0
            ref = O.gaussian_kernel1d(k, sigma, name)
            got = np.array(buf[:k], dtype=np.float32)
            if name == "float32":
                np.testing.assert_allclose(got, ref, rtol=3e-7, atol=0)
            else:
                np.testing.assert_array_equal(got, ref)


def test_argument_validation_reports_errors():
    from alg_b200 import _lib

    L = _lib.lib()
    assert L.alg_lowpass_gaussian(None, None, 1, 8, 8, 3, 1.0, 0, None) != 0
    assert b"null" in L.alg_last_error()
    assert L.alg_gaussian_kernel1d(99, 1.0, 0, (ctypes.c_float * 4)()) != 0


def test_no_cpu_fallback():
    from alg_b200 import lowpass

    x = torch.zeros(1, 1, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lowpass.apply_low_pass_filter(x, "down_up", 1.0, 3, 0.5)
    # early exits never touch the device and return the same object (lp_utils.py:23-28)
    assert lowpass.apply_low_pass_filter(x, "none", 1.0, 3, 0.5) is x
    assert lowpass.apply_low_pass_filter(x, "down_up", 1.0, 3, 1.0) is x
    assert lowpass.apply_low_pass_filter(x, "gaussian_blur", 0, 3, 0.5) is x


def test_product_code_never_imports_oracle():
    bad = []
    for base in ("alg_b200", "."):
        for fn in os.listdir(os.path.join(ROOT, base)):
            if fn.endswith(".py") and fn not in ("bench.py", "__graft_entry__.py"):
                src = open(os.path.join(ROOT, base, fn)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M):
                    bad.append(fn)
    assert not bad, bad


def test_strength_and_bucketing_match_reference(golden_dir):
    from alg_b200 import lowpass
    import contextlib, io

    rows = json.load(open(os.path.join(golden_dir, "lp_strength.json")))
    with contextlib.redirect_stdout(io.StringIO()):
        for r in rows:
            assert lowpass.get_lp_strength(*r[:9]) == r[9]

    class Img:
        def __init__(self, w, h):
            self.size = (w, h)

    for res, w, h, th, tw in json.load(open(os.path.join(golden_dir, "hunyuan_size.json"))):
        assert lowpass.get_hunyuan_video_size(res, Img(w, h)) == (th, tw)
    assert lowpass.get_hunyuan_video_size("360p", Img(832, 480)) == (352, 608)
    assert lowpass.get_hunyuan_video_size("720p", Img(832, 480)) == (736, 1248)


def test_unipc_timesteps_known_answer():
    from alg_b200 import schedulers as S

    s = S.UniPCMultistepScheduler(flow_shift=5.0)
    s.set_timesteps(50)
    ts = s.timesteps.tolist()
    assert ts[:6] == [999, 995, 991, 987, 982, 978] and ts[-4:] == [302, 241, 172, 92]
    assert abs(float(s.sigmas[0]) - 0.99980) < 1e-5 and abs(float(s.sigmas[49]) - 0.092507) < 1e-6
    s3 = S.UniPCMultistepScheduler.from_config(s.config, flow_shift=3.0)
    s3.set_timesteps(50)
    assert s3.timesteps.tolist()[:6] == [999, 992, 985, 978, 971, 963]


def test_scheduler_host_scalars_match_oracle():
    """Same coefficients as the restated diffusers scalar math (they feed the fused kernel)."""
    from alg_b200 import schedulers as S
    from oracle import sched_oracle as O

    s = S.UniPCMultistepScheduler(flow_shift=5.0)
    o = O.UniPCOracle(flow_shift=5.0)
    for n in (2, 5, 50):
        s.set_timesteps(n)
        o.set_timesteps(n)
        assert torch.equal(s.timesteps, o.timesteps) and torch.equal(s.sigmas, o.sigmas)
    d, od = S.CogVideoXDDIMScheduler(), O.CogDDIMOracle()
    d.set_timesteps(50)
    od.set_timesteps(50)
    assert d.timesteps.tolist() == od.timesteps.tolist() and d.timesteps.tolist()[:3] == [999, 979, 959]
    for t in (999, 499, 19):
        a_t, b_t, a, b = od.coeffs(t)
        assert d._coeffs(t) == (float(a_t ** 0.5), float(b_t ** 0.5), float(a), float(b))
    e, oe = S.FlowMatchEulerDiscreteScheduler(shift=7.0), O.FlowEulerOracle(shift=7.0)
    sig = torch.linspace(1, 0, 31)[:-1].numpy()
    e.set_timesteps(sigmas=sig)
    oe.set_timesteps(30, sigmas=sig)
    assert torch.equal(e.sigmas, oe.sigmas) and torch.equal(e.timesteps, oe.timesteps)
    # run.py:82 passes flow_shift= to a scheduler whose parameter is named shift: ignored like in diffusers
    e2 = S.FlowMatchEulerDiscreteScheduler.from_config(e.config, flow_shift=17.0, invert_sigmas=False)
    assert e2.config.shift == 7.0 and e2.ignored_config_keys == ["flow_shift"]


def test_dpm_host_scalars_match_oracle():
    """CogVideoXDPMScheduler coefficients (host fp64 -> fp32) against the oracle's, including the zero-terminal-SNR first
    step (alpha = 0: h = inf, mult finite) and the final step onto alpha = 1."""
    from alg_b200 import schedulers as S
    from oracle import sched_oracle as O
    e, o = S.CogVideoXDPMScheduler(), O.CogDPMOracle()
    for n in (2, 10, 50):
        e.set_timesteps(n)
        o.set_timesteps(n)
        ts = o.timesteps.tolist()
        assert e.timesteps.tolist() == ts
        for i, t in enumerate(ts):
            tb = ts[i - 1] if i > 0 else None
            k = e._dpm_coeffs(t, tb)
            a_t, prev_t, mult, mn = o.dpm_coeffs(t, tb)
            got = [k["m0"], k["m1"]] + ([k["m2"], k["m3"]] if tb is not None else [])
            used = got[:2] if (tb is None or prev_t < 0) else got  # mult[2:] feed the second-order update only
            assert k["prev_t"] == prev_t and all(np.isfinite(used)) and np.isfinite(k["mn"])
            np.testing.assert_array_equal(got, [float(m) for m in mult])  # NaN == NaN here (unused last-step r = 0)
            assert k["mn"] == float(mn) and k["sa"] == float(a_t ** 0.5)


def test_double_float_rope_equals_complex128_after_bf16_rounding():
    """Host model of dd_dot2 (alg_b200/csrc/dit_kernels.cu): re*cos - im*sin evaluated with fp32 (hi, lo) tables, exact FMA
    residuals and TwoSum, one final fp32 rounding -- after the cast to bf16 it equals the complex128 product diffusers
    computes (wan DiT RoPE), where a plain fp32 FMA evaluation does not."""
    import torch
    rng = np.random.default_rng(0)
    n = 2_000_000
    f32, f64 = np.float32, np.float64

    def bf16(x):
        return torch.from_numpy(x).to(torch.bfloat16).float().numpy()

    def to_bf16_bits(x):
        return torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy()

    def fma(a, b, c):  # fp32 FMA: the product of two fp32 is exact in fp64
        return (a.astype(f64) * b.astype(f64) + c.astype(f64)).astype(f32)

    re, im = bf16((rng.standard_normal(n) * 2).astype(f32)), bf16((rng.standard_normal(n) * 2).astype(f32))
    ang = rng.uniform(0, 1000, n)
    c, s = np.cos(ang), np.sin(ang)
    ref_re = (re.astype(f64) * c - im.astype(f64) * s).astype(f32)
    ref_im = (re.astype(f64) * s + im.astype(f64) * c).astype(f32)
    ch, sh = c.astype(f32), s.astype(f32)
    cl, sl = (c - ch.astype(f64)).astype(f32), (s - sh.astype(f64)).astype(f32)

    def dd(a, xh, xl, b, yh, yl):
        p1 = (a * xh).astype(f32)
        e1 = fma(a, xh, -p1)
        p2 = (b * yh).astype(f32)
        e2 = fma(b, yh, -p2)
        ss = (p1 + p2).astype(f32)
        bb = (ss - p1).astype(f32)
        err = ((p1 - (ss - bb).astype(f32)).astype(f32) + (p2 - bb).astype(f32)).astype(f32)
        low = (((e1 + e2).astype(f32) + err).astype(f32) + fma(a, xl, (b * yl).astype(f32))).astype(f32)
        return (ss + low).astype(f32)

    got_re, got_im = dd(re, ch, cl, -im, sh, sl), dd(re, sh, sl, im, ch, cl)
    assert np.array_equal(to_bf16_bits(got_re), to_bf16_bits(ref_re))
    assert np.array_equal(to_bf16_bits(got_im), to_bf16_bits(ref_im))
    assert (got_re != ref_re).mean() < 1e-5  # even the fp32 values agree almost everywhere
    plain = fma(re, ch, -(im * sh).astype(f32))
    assert (to_bf16_bits(plain) != to_bf16_bits(ref_re)).sum() > 0


def test_unipc_coefficients_against_fp64_closed_forms():
    """The scheduler's per-step scalars (computed once in set_timesteps with the 0-dim fp32 torch op chain diffusers uses)
    against an INDEPENDENT derivation: fp64, no logs / expm1 for the exponential terms -- with flow sigmas
    exp(-h) = (sigma_t alpha_s0) / (alpha_t sigma_s0), so alpha_t * h_phi_1 = alpha_t * B_h = sigma_t alpha_s0 / sigma_s0 - alpha_t
    -- and the 2 x 2 corrector system R rho = b solved in closed form (product and oracle shared one text)."""
    import math

    import numpy as np

    from alg_b200.schedulers import UniPCMultistepScheduler
    for n, shift in ((50, 5.0), (30, 3.0), (8, 1.0)):
        s = UniPCMultistepScheduler(flow_shift=shift)
        s.set_timesteps(n)
        sig = s.sigmas.double().numpy()  # the fp32 schedule values ARE the definition; everything after is fp64 here
        lam = lambda i: math.log((1 - sig[i]) / sig[i]) if sig[i] > 0 else math.inf
        seen = 0
        for (i_t, i_s0, order, corrector), k in s._coef.items():
            st, s0 = sig[i_t], sig[i_s0]
            at, a0 = 1 - st, 1 - s0
            want_ratio = st / s0
            want_a = st * a0 / s0 - at  # alpha_t * expm1(-h)
            assert abs(k["ratio"] - want_ratio) <= 2e-6 * max(1.0, abs(want_ratio)), (i_t, k["ratio"], want_ratio)
            assert abs(k["a"] - want_a) <= 4e-6 * max(1.0, abs(want_a)), (i_t, k["a"], want_a)
            assert abs(k["b"] - want_a) <= 4e-6 * max(1.0, abs(want_a))  # bh2: B_h = expm1(-h) too
            if order == 2:
                lt, l0, lp = lam(i_t), lam(i_s0), lam(i_s0 - 1)
                if math.isinf(lt):  # last step: sigma_t = 0, h = inf, rk = 0 (diffusers divides by it: inf; the kernel multiplies by 1/rk)
                    continue
                h = lt - l0
                rk = (lp - l0) / h
                assert abs(k["rk_inv"] * rk - 1) < 2e-5, (i_t, k["rk_inv"], 1 / rk)
                if corrector:
                    hh = -h
                    e1 = math.expm1(hh)
                    k1 = e1 / hh - 1
                    k2 = k1 / hh - 0.5
                    b0, b1 = k1 / e1, 2 * k2 / e1
                    rho0 = (b0 - b1) / (1 - rk)
                    rho1 = b0 - rho0
                    ok0 = abs(k["rho0"] - rho0) < 5e-5 * max(1.0, abs(rho0))
                    ok1 = abs(k["rho_last"] - rho1) < 5e-5 * max(1.0, abs(rho1))
                    assert ok0 and ok1, (i_t, k["rho0"], rho0, k["rho_last"], rho1)
                else:
                    assert k["rho0"] == 0.5
            else:
                assert k["rho0"] == 0.5 and k["rho_last"] == 0.5
            seen += 1
        assert seen >= 2 * n - 3
    # schedule definition itself: sigma_i = shift * s / (1 + (shift - 1) s), s = linspace(1, 1/1000, n + 1)[:-1] reversed
    s = UniPCMultistepScheduler(flow_shift=5.0)
    s.set_timesteps(50)
    base = 1 - np.linspace(1, 1 / 1000, 51)
    want = np.flip(5.0 * base / (1 + 4.0 * base))[:-1]
    assert np.allclose(s.sigmas.numpy()[:-1], want.astype(np.float32)) and float(s.sigmas[-1]) == 0.0
    assert s.timesteps.tolist() == [int(v) for v in (want * 1000).astype(np.int64)]


def test_ddim_coefficients_against_fp64_closed_forms():
    """CogVideoXDDIMScheduler._coeffs vs the v-prediction DDIM update written out in fp64 from alphas_cumprod:
    x_prev = sqrt(a_prev) x0 + sqrt(1 - a_prev) eps with x0 = sqrt(a_t) x - sqrt(1 - a_t) v, eps = sqrt(a_t) v + sqrt(1 - a_t) x,
    which must equal the scheduler's  a * x + b * x0  form for every step."""
    import math

    from alg_b200.schedulers import CogVideoXDDIMScheduler
    s = CogVideoXDDIMScheduler()
    s.set_timesteps(50)
    ac = s.alphas_cumprod.double()
    for t in s.timesteps.tolist():
        sa, sb, a, b = s._coeffs(int(t))
        prev = t - 1000 // 50
        a_t = float(ac[t])
        a_p = float(ac[prev]) if prev >= 0 else 1.0
        assert abs(sa - math.sqrt(a_t)) < 1e-6 and abs(sb - math.sqrt(1 - a_t)) < 1e-6
        x, v = 0.37, -1.21
        x0 = math.sqrt(a_t) * x - math.sqrt(1 - a_t) * v
        eps = math.sqrt(a_t) * v + math.sqrt(1 - a_t) * x
        want = math.sqrt(a_p) * x0 + math.sqrt(1 - a_p) * eps
        got = a * x + b * (sa * x - sb * v)
        assert abs(got - want) < 2e-6 * max(1.0, abs(want)), (t, got, want)


def test_pipeline_module_helpers_and_qkv_fusion_bookkeeping(caplog):
    """Reference module-level helpers (wan:97-112) and the Cog pipeline's fuse/unfuse bookkeeping (cog:527-539)."""
    import logging

    import pipeline_cogvideox_image2video_lowpass as cog
    import pipeline_wan_image2video_lowpass as wan

    assert wan.basic_clean("  a &amp;amp; b  ") == "a & b"
    assert wan.whitespace_clean(" a \n\t b   c ") == "a b c"
    assert wan.prompt_clean("  x &amp;lt;  \n y ") == "x < y"
    assert cog.get_resize_crop_region_for_grid((30, 45), 45, 30) == ((0, 0), (30, 45))

    class _T:
        fused = 0

        def fuse_qkv_projections(self):
            self.fused += 1

        def unfuse_qkv_projections(self):
            self.fused -= 1

    pipe = cog.CogVideoXImageToVideoPipeline.__new__(cog.CogVideoXImageToVideoPipeline)
    pipe.transformer = _T()
    with caplog.at_level(logging.WARNING):
        pipe.unfuse_qkv_projections()
    assert "not initially fused" in caplog.text and pipe.transformer.fused == 0
    pipe.fuse_qkv_projections()
    assert pipe.fusing_transformer and pipe.transformer.fused == 1
    pipe.unfuse_qkv_projections()
    assert not pipe.fusing_transformer and pipe.transformer.fused == 0
