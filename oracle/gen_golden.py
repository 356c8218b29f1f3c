"""Generate tests/golden/lp_*.npz by running the REAL reference ``lp_utils.py``.

Run in the build container only (``/root/reference`` is not on the GPU box):

    python oracle/gen_golden.py

The reference module is imported from where it lies (never copied).  The
fixtures pin ``oracle/lp_oracle.py`` (tests/test_oracle_lp.py) and are what the
``-m gpu`` parity tests compare the CUDA kernels against.
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("ALG_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    sys.path.insert(0, REF)
    import lp_utils  # the reference, unmodified

    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(1234)
    torch.set_num_threads(1)

    # ---- strength schedule (lp_utils.py:63-111) -------------------------------
    rows = []
    for kind in ("linear", "interval", "exponential", "none", "bogus"):
        for total in (1, 2, 30, 50):
            for (i0, i1, ls, le, lt, er) in (
                (0.0, 0.2, 1.0, 0.0, 0.5, 10.0),
                (0.0, 0.04, 1.0, 0.0, 0.5, 10.0),
                (0.1, 0.3, 0.8, 0.2, 0.25, 3.0),
                (0.0, 0.05, 1.0, 0.5, 0.0, -2.0),
            ):
                for step in sorted({0, 1, 2, 5, 10, total // 2, max(total - 1, 0)}):
                    if step >= max(total, 1):
                        continue
                    import contextlib, io

                    with contextlib.redirect_stdout(io.StringIO()):
                        v = lp_utils.get_lp_strength(step, total, kind, i0, i1, ls, le, lt, er)
                    rows.append([step, total, kind, i0, i1, ls, le, lt, er, float(v)])
    with open(os.path.join(OUT, "lp_strength.json"), "w") as f:
        json.dump(rows, f)

    # ---- down_up (lp_utils.py:49-54), fp32 on CPU (quirk q16: no bf16 on CPU) --
    cases = {}
    specs = [
        ("wan_cfg2", (1, 3, 2, 60, 104), 0.4),      # [B,C,K,H,W] 5-D, cfg 2 plane geometry
        ("cog_cfg1", (1, 2, 3, 60, 90), 0.25),      # 90*0.25 = 22.5 -> 22 (banker's)
        ("hy_cfg4", (1, 2, 1, 90, 160), 0.625),
        ("small_4d", (2, 3, 30, 45), 0.25),         # -> (8, 11)
        ("odd", (1, 2, 17, 23), 0.5),
        ("third", (1, 1, 31, 64), 1.0 / 3.0),
        ("tiny_to_1", (1, 1, 5, 7), 0.05),          # h1 = w1 = 1
        ("up_identity_like", (1, 1, 16, 16), 0.95),
        ("modulated", (1, 2, 2, 60, 104), 1.0 - (1.0 - 0.4) * 0.3604),
    ]
    for name, shape, f in specs:
        x = torch.randn(shape, dtype=torch.float32)
        y = lp_utils.apply_low_pass_filter(x, "down_up", 0.0, 0.0, f)
        cases[name + "_x"] = x.numpy()
        cases[name + "_y"] = y.numpy()
        cases[name + "_f"] = np.float64(f)
    # constant planes are preserved (Wan's 4 mask channels, quirk q5)
    x = torch.ones(1, 1, 1, 60, 104)
    cases["const_y"] = lp_utils.apply_low_pass_filter(x, "down_up", 0.0, 0.0, 0.4).numpy()
    np.savez_compressed(os.path.join(OUT, "lp_down_up.npz"), **cases)

    # ---- gaussian_blur (lp_utils.py:40-47), fp32 and bf16 ----------------------
    cases = {}
    specs = [
        ("cog_pix_f32", (1, 3, 480, 64), torch.float32, 15.0, 0.02734375),   # k = 13 from H = 480
        ("cog_pix_bf16", (1, 3, 480, 64), torch.bfloat16, 15.0, 0.02734375),
        ("h512_bf16", (1, 1, 512, 40), torch.bfloat16, 7.5, 0.02734375),     # k = 15
        ("int_k_f32", (2, 2, 33, 47), torch.float32, 2.0, 7),
        ("even_int_k_f32", (1, 2, 3, 20, 24), torch.float32, 1.25, 4),       # -> 5, 5-D
        ("latent_k1_f32", (1, 2, 2, 60, 104), torch.float32, 15.0, 0.02734375),  # k = 1
        ("small_sigma_bf16", (1, 2, 40, 40), torch.bfloat16, 0.6, 9),
        ("sched_k_f32", (1, 1, 480, 32), torch.float32, 15.0 * 0.5, 0.02734375 * 0.5),  # k*s float -> 6 -> 7
    ]
    for name, shape, dt, sigma, k in specs:
        x = torch.randn(shape, dtype=torch.float32).to(dt)
        y = lp_utils.apply_low_pass_filter(x, "gaussian_blur", sigma, k, 0.0)
        cases[name + "_x"] = x.float().numpy()
        cases[name + "_y"] = y.float().numpy()
        cases[name + "_sigma"] = np.float64(sigma)
        cases[name + "_k"] = np.float64(k) if isinstance(k, float) else np.int64(k)
        cases[name + "_dtype"] = np.array(str(dt).replace("torch.", ""))
    np.savez_compressed(os.path.join(OUT, "lp_gaussian.npz"), **cases)

    # ---- early exits return the same object (lp_utils.py:23-28) ----------------
    x = torch.randn(1, 2, 8, 8)
    assert lp_utils.apply_low_pass_filter(x, "none", 1.0, 3, 0.5) is x
    assert lp_utils.apply_low_pass_filter(x, "down_up", 1.0, 3, 1.0) is x
    assert lp_utils.apply_low_pass_filter(x, "gaussian_blur", 0, 3, 0.5) is x

    # ---- Hunyuan bucketing (lp_utils.py:113-189) -------------------------------
    class Img:  # only ``.size`` (W, H) is read
        def __init__(self, w, h):
            self.size = (w, h)

    rows = []
    for res in ("360p", "540p", "720p"):
        for (w, h) in ((832, 480), (887, 512), (480, 832), (512, 512), (1920, 1080), (300, 1000)):
            th, tw = lp_utils.get_hunyuan_video_size(res, Img(w, h))
            rows.append([res, w, h, int(th), int(tw)])
    with open(os.path.join(OUT, "hunyuan_size.json"), "w") as f:
        json.dump(rows, f)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
