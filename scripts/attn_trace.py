"""Per-CTA timeline of the attention kernel (clock64 stamps, see ATTN_TRACE in csrc/attention.cu).

    python scripts/attn_trace.py build            # here (no GPU): alg_b200/libalg_b200_trace.so, compiled with -DALG_ATTN_TRACE
    python scripts/attn_trace.py run [shape ...]  # on the GPU box: medians of the phase durations per shape

Shapes: wan_cross_text, wan_cross_text_acc (accumulating epilogue), wan_cross_image, wan_self (8 heads).
The product library is untouched: the probe is compiled out of libalg_b200.so.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "alg_b200", "csrc")
TRACE_LIB = os.path.join(ROOT, "alg_b200", "libalg_b200_trace.so")
SRCS = "capi lowpass cfg_sched gemm attention dit_kernels dit_ops vae_ops vae_f32_ops encoder_ops wan_engine".split()


def build():
    out = os.path.join(CSRC, "build_trace")
    os.makedirs(out, exist_ok=True)
    flags = ("-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DALG_ATTN_TRACE").split()
    procs = [subprocess.Popen(["nvcc", *flags, "-c", os.path.join(CSRC, f + ".cu"), "-o", os.path.join(out, f + ".o")]) for f in SRCS]
    assert all(p.wait() == 0 for p in procs)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", TRACE_LIB,
                           *[os.path.join(out, f + ".o") for f in SRCS], "-cudart", "static"])
    print("built", TRACE_LIB)


SHAPES = {"wan_cross_text": (3, 40, 128, 32760, 512, False), "wan_cross_text_acc": (3, 40, 128, 32760, 512, True),
          "wan_cross_image": (3, 40, 128, 32760, 257, False), "wan_self": (1, 8, 128, 32760, 32760, False)}
PHASES = (("entry -> setup done (barriers, TMEM alloc, sync)", 0, 1), ("setup -> Q landed (issuer)", 1, 4),
          ("Q landed -> S(0) issued", 4, 5), ("setup -> softmax sees S(0)", 1, 7), ("softmax step 0 (S(0) seen -> P(0) published)", 7, 8),
          ("P(0) published -> softmax loop done", 8, 9), ("softmax loop done -> last PV complete", 9, 10),
          ("epilogue (O / l -> global)", 10, 11), ("  epilogue: o_full seen -> first tcgen05.ld back", 10, 14),
          ("  epilogue: first ld -> rows staged in smem", 14, 13), ("  epilogue: staged -> stores issued", 13, 11), ("epilogue done -> CTA past final sync", 11, 12), ("CTA lifetime", 0, 12),
          ("issuer: S(0) issued -> last PV issued", 5, 6))


def run(names):
    import ctypes as C

    import numpy as np
    import torch
    from alg_b200 import _lib
    _lib.LIB_PATH = TRACE_LIB
    L = _lib.lib()
    L.alg_attention_trace_buffer.restype = C.c_int
    L.alg_attention_trace_buffer.argtypes = [C.c_void_p]
    from alg_b200 import ops
    for name in names:
        B, H, D, Nq, N, acc = SHAPES[name]
        q = torch.randn(B, Nq, H, D, device="cuda").bfloat16()
        k = torch.randn(B, N, H, D, device="cuda").bfloat16()
        vt = torch.randn(B, H, D, (N + 7) // 8 * 8, device="cuda").bfloat16()
        o = torch.zeros_like(q)
        tiles = 1 if N <= 1024 else 2
        n_cta = B * H * ((Nq + tiles * 128 - 1) // (tiles * 128))
        buf = torch.zeros(n_cta, 32, dtype=torch.int64, device="cuda")
        for _ in range(2):
            ops.attention(q, k, vt, n_kv=N, out=o, accumulate=acc)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        assert L.alg_attention_trace_buffer(C.c_void_p(buf.data_ptr())) == 0
        torch.cuda.synchronize()
        e0.record()
        ops.attention(q, k, vt, n_kv=N, out=o, accumulate=acc)
        e1.record()
        torch.cuda.synchronize()
        assert L.alg_attention_trace_buffer(None) == 0
        ms = e0.elapsed_time(e1)
        t = buf.cpu().numpy()
        per_sm = np.bincount(t[:, 2].astype(np.int64), minlength=148)
        # kernel span per SM in that SM's clock: first entry to last exit
        span = np.array([t[t[:, 2] == s, 12].max() - t[t[:, 2] == s, 0].min() for s in range(148) if per_sm[s]])
        life = t[:, 12] - t[:, 0]
        print(f"== {name}: B={B} H={H} Nq={Nq} Nkv={N} accumulate={acc}: {ms:.3f} ms, {4 * B * H * Nq * N * D / ms / 1e9:.0f} TFLOP/s, "
              f"{n_cta} CTAs ({per_sm.mean():.1f} per SM), SM span median {np.median(span):.0f} clk -> "
              f"{np.median(span) / per_sm.mean():.0f} clk per CTA slot; sum of CTA lifetimes / span = {life.sum() / span.sum():.2f} CTAs resident")
        for label, a, b in PHASES:
            d = (t[:, b] - t[:, a]).astype(np.float64)
            print(f"   {label:52s} median {np.median(d):8.0f}   p10 {np.percentile(d, 10):8.0f}   p90 {np.percentile(d, 90):8.0f} clk")
        del q, k, vt, o, buf


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        run(sys.argv[2:] or list(SHAPES))
