"""Native ``AutoencoderKLWan`` (alg_b200/vae_wan.py: whole-clip evaluation on the fp32-split tensor-core path) against
``oracle/wan_vae_oracle.py`` (diffusers' CHUNKED feature-cache evaluation restated with F.conv3d / F.normalize / SDPA; parity
unpinned, diffusers absent).  The oracle runs in float64 on the GPU; tolerance 1e-4 relative L2 on the moments / the decoded
clip (the bf16 3-term split carries ~2^-16 per product; TF32, which PyTorch's own fp32 convolutions use by default, would be 1e-3)."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TINY = dict(base_dim=16, z_dim=4, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[], temperal_downsample=[False, True, True])


def _pair(cfg, seed=1):
    from alg_b200.vae_wan import AutoencoderKLWan
    from oracle import wan_vae_oracle as V
    ocfg = dict(V.WAN21_VAE, **cfg)
    sd = V.make_weights(ocfg, seed=seed, device="cuda", dtype=torch.float32)
    vae = AutoencoderKLWan(**cfg).load_state_dict({k: v.clone() for k, v in sd.items()})
    return vae, sd, ocfg


def test_im2col_split3_geometries_against_unfold():
    """The fused gather + split alone: causal zero padding, stride 2 with bottom/right zero pad, x2 nearest upsample, t_min,
    frame ranges, the scalar (C = 3) path; hi + lo must reproduce the fp32 patch to 2^-16."""
    import ctypes as C
    import torch.nn.functional as F
    from alg_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(0)

    def run(T, H, W, Cc, k, stride, pad, up, t_min, to0, To, Ho, Wo):
        x = torch.randn(T, H, W, Cc, generator=g, device="cuda")
        K = k[0] * k[1] * k[2] * Cc
        ld = (K + 7) // 8 * 8
        cols = torch.empty(To * Ho * Wo, 3 * ld, device="cuda", dtype=torch.bfloat16)
        p = _lib.Im2colF32()
        p.x, p.cols, p.T, p.H, p.W, p.C = x.data_ptr(), cols.data_ptr(), T, H, W, Cc
        p.kt, p.kh, p.kw, p.st, p.sh, p.sw = *k, *stride
        p.pad_t, p.pad_top, p.pad_left, p.To, p.Ho, p.Wo, p.to0, p.up, p.t_min, p.ld = *pad, To, Ho, Wo, to0, up, t_min, ld
        p.replicate, p.tdup = 0, 1
        _lib.check(_lib.lib().alg_im2col_split3_f32(C.byref(p), _lib.stream_ptr(x.device)))
        torch.cuda.synchronize()
        # reference: explicit loops over taps on the (upsampled, zero-padded) clip
        xl = x.repeat_interleave(up, 1).repeat_interleave(up, 2) if up > 1 else x
        xl = xl.clone()
        xl[:t_min] = 0
        HL, WL = xl.shape[1:3]
        want = torch.zeros(To, Ho, Wo, k[0], k[1], k[2], Cc, device="cuda")
        for it in range(k[0]):
            for ih in range(k[1]):
                for iw in range(k[2]):
                    for to in range(To):
                        t = (to0 + to) * stride[0] + it - pad[0]
                        if t < 0:
                            continue
                        ys = torch.arange(Ho, device="cuda") * stride[1] + ih - pad[1]
                        xs = torch.arange(Wo, device="cuda") * stride[2] + iw - pad[2]
                        oky, okx = (ys >= 0) & (ys < HL), (xs >= 0) & (xs < WL)
                        patch = xl[t][ys.clamp(0, HL - 1)][:, xs.clamp(0, WL - 1)] * (oky[:, None, None] & okx[None, :, None])
                        want[to, :, :, it, ih, iw] = patch
        want = want.reshape(To * Ho * Wo, K)
        hi, hi2, lo = cols[:, :K].float(), cols[:, ld:ld + K].float(), cols[:, 2 * ld:2 * ld + K].float()
        assert torch.equal(hi, hi2) and torch.equal(hi, want.bfloat16().float())
        assert (hi + lo - want).abs().max() <= want.abs().max() * 2 ** -15
        if ld > K:
            assert bool((cols[:, K:ld] == 0).all()) and bool((cols[:, 2 * ld + K:] == 0).all())

    run(5, 6, 10, 8, (3, 3, 3), (1, 1, 1), (2, 1, 1), 1, 0, 0, 5, 6, 10)        # WanCausalConv3d 3x3x3
    run(5, 6, 10, 8, (3, 3, 3), (1, 1, 1), (2, 1, 1), 1, 0, 2, 2, 6, 10)        # a frame chunk of it
    run(3, 7, 10, 16, (1, 3, 3), (1, 2, 2), (0, 0, 0), 1, 0, 0, 3, 3, 5)        # ZeroPad2d((0,1,0,1)) + stride 2 (odd height)
    run(3, 4, 6, 12, (1, 3, 3), (1, 1, 1), (0, 1, 1), 2, 0, 0, 3, 8, 12)        # nearest x2 upsample fused
    run(7, 4, 6, 8, (3, 1, 1), (1, 1, 1), (2, 0, 0), 1, 1, 1, 6, 4, 6)          # upsample3d time_conv: frame 0 reads as zero
    run(9, 4, 6, 8, (3, 1, 1), (2, 1, 1), (2, 0, 0), 1, 0, 1, 4, 4, 6)          # downsample3d time_conv, outputs k >= 1
    run(5, 6, 10, 3, (3, 3, 3), (1, 1, 1), (2, 1, 1), 1, 0, 0, 5, 6, 10)        # RGB input: scalar path, K = 81 -> ld 88


@pytest.mark.parametrize("T,h,w", [(4, 4, 6), (1, 6, 4), (2, 5, 6)])
def test_decode_matches_chunked_oracle(T, h, w):
    from oracle import wan_vae_oracle as V
    vae, sd, ocfg = _pair(TINY)
    z = torch.randn(1, TINY["z_dim"], T, h, w, generator=torch.Generator(device="cuda").manual_seed(2), device="cuda")
    out = vae.decode(z).sample
    ref = V.decode(z.double(), {k: v.double() for k, v in sd.items()}, ocfg, torch.float64)
    assert out.shape == ref.shape == (1, 3, 4 * T - 3, 8 * h, 8 * w) and out.dtype == torch.float32
    assert float(out.abs().max()) <= 1.0 and rel_l2(out, ref) < 1e-4, rel_l2(out, ref)
    assert vae.decode(z, return_dict=False)[0].shape == out.shape


@pytest.mark.parametrize("T,H,W", [(9, 32, 48), (1, 48, 32), (5, 40, 48)])
def test_encode_matches_chunked_oracle(T, H, W):
    from oracle import wan_vae_oracle as V
    vae, sd, ocfg = _pair(dict(TINY, z_dim=8), seed=3)
    x = torch.rand(1, 3, T, H, W, generator=torch.Generator(device="cuda").manual_seed(4), device="cuda") * 2 - 1
    dist = vae.encode(x).latent_dist
    ref = V.encode_moments(x.double(), {k: v.double() for k, v in sd.items()}, ocfg, torch.float64)
    assert dist.parameters.shape == ref.shape == (1, 16, 1 + (T - 1) // 4, H // 8, W // 8)
    assert rel_l2(dist.parameters, ref) < 1e-4, rel_l2(dist.parameters, ref)
    assert torch.equal(dist.mode(), dist.parameters[:, :8])
    s1 = dist.sample(torch.Generator(device="cuda").manual_seed(5))
    s2 = dist.sample(torch.Generator(device="cuda").manual_seed(5))
    assert torch.equal(s1, s2) and s1.shape == dist.mean.shape


def test_batch_of_two_and_surface():
    """wan:180-181 reads ``vae.temperal_downsample`` off the module; config carries z_dim / latents_mean / latents_std;
    a batch is one clip after another; CPU tensors and wrong channel counts are refused."""
    from oracle import wan_vae_oracle as V
    vae, sd, ocfg = _pair(TINY)
    assert vae.temperal_downsample == [False, True, True] and vae.config.z_dim == 4 and vae.dtype == torch.float32
    assert len(vae.config.latents_mean) == 16
    z = torch.randn(2, 4, 2, 4, 4, generator=torch.Generator(device="cuda").manual_seed(6), device="cuda")
    out = vae.decode(z).sample
    for b in range(2):
        assert torch.equal(out[b:b + 1], vae.decode(z[b:b + 1]).sample)
    with pytest.raises(RuntimeError, match="CUDA"):
        vae.decode(z.cpu())
    with pytest.raises(ValueError, match="expected"):
        vae.decode(z[:, :3])
    assert set(vae.state_dict()) == set(sd)


def test_from_pretrained_snapshot_and_wan_pipeline_condition(tmp_path):
    """run.py:51-55 + wan:372-449: a snapshot's ``vae/`` folder loads into the native VAE; the pipeline's ``prepare_latents``
    (image + zero frames -> encode -> argmax -> normalise, mask in front) matches the same code on the oracle VAE."""
    from types import SimpleNamespace
    from alg_b200 import checkpoint
    from alg_b200.schedulers import UniPCMultistepScheduler
    from alg_b200.vae_wan import AutoencoderKLWan
    from oracle import wan_vae_oracle as V
    from pipeline_wan_image2video_lowpass import WanImageToVideoPipeline
    cfg = dict(V.WAN21_VAE, base_dim=16)
    sd = V.make_weights(cfg, seed=7, device="cuda")
    snap = str(tmp_path / "snap")
    checkpoint.save_component(snap, "vae", dict(cfg, _class_name="AutoencoderKLWan"), sd)
    vae = AutoencoderKLWan.from_pretrained(snap)
    assert vae.config.base_dim == 16 and vae.config.z_dim == 16

    class OracleVAE:
        dtype = torch.float32
        temperal_downsample = cfg["temperal_downsample"]
        config = SimpleNamespace(**cfg)

        def encode(self, x):
            m = V.encode_moments(x.double(), {k: v.double() for k, v in sd.items()}, cfg, torch.float64).float()
            return SimpleNamespace(latent_dist=SimpleNamespace(mode=lambda: m[:, :16], sample=lambda generator=None: m[:, :16]))

    dummy = SimpleNamespace(config=SimpleNamespace(patch_size=(1, 2, 2)), dtype=torch.bfloat16, to=lambda *a, **k: None)
    res = []
    for v in (vae, OracleVAE()):
        pipe = WanImageToVideoPipeline(tokenizer=None, text_encoder=None, image_encoder=None, image_processor=None, transformer=dummy,
                                       vae=v, scheduler=UniPCMultistepScheduler())
        image = torch.rand(1, 3, 32, 48, generator=torch.Generator().manual_seed(8)) * 2 - 1
        lat, cond = pipe.prepare_latents(image, 1, 16, 32, 48, 9, torch.float32, torch.device("cuda"),
                                         torch.Generator(device="cuda").manual_seed(9), None)
        res.append((lat, cond))
    assert torch.equal(res[0][0], res[1][0]) and res[0][1].shape == res[1][1].shape == (1, 20, 3, 4, 6)
    assert torch.equal(res[0][1][:, :4], res[1][1][:, :4])  # the first-frame mask
    assert rel_l2(res[0][1][:, 4:], res[1][1][:, 4:]) < 1e-4


def test_wan_pipeline_end_to_end_all_native_vs_upstream_modules():
    """run.py's whole path on tiny shapes -- image + prompt -> tokenizer -> UMT5, CLIP, VAE encode of the condition clip,
    denoise loop with ALG, VAE decode, post-processing to frames: once with every network native, once with the transformers
    modules (same weights) and the oracle VAE around the same native DiT.  Frames must agree closely and be valid video."""
    from types import SimpleNamespace
    import numpy as np
    from transformers import CLIPVisionConfig, CLIPVisionModel as HFCLIP
    from alg_b200 import encoders
    from alg_b200.pipeline_utils import SyntheticImageProcessor, SyntheticTokenizer
    from alg_b200.schedulers import UniPCMultistepScheduler
    from alg_b200.vae_wan import AutoencoderKLWan
    from oracle import wan_vae_oracle as V
    from pipeline_wan_image2video_lowpass import WanImageToVideoPipeline
    import __graft_entry__ as G
    from test_gpu_encoders import _umt5
    cfg, dit, _, alg = G.tiny_problem("cuda")
    base, hf_t = _umt5("UMT5EncoderModel", dict(vocab_size=4096, d_model=64, d_kv=16, d_ff=128))
    hf_t = hf_t.to(torch.bfloat16).cuda()
    ccfg = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=2, image_size=224, patch_size=28)
    torch.manual_seed(5)
    hf_c = HFCLIP(CLIPVisionConfig(**ccfg)).eval().float().cuda()
    text = encoders.UMT5EncoderModel(**base).load_state_dict({k: v.clone() for k, v in hf_t.state_dict().items()})
    clip = encoders.CLIPVisionModel(**ccfg).load_state_dict({k: v.clone() for k, v in hf_c.state_dict().items()})
    vcfg = dict(V.WAN21_VAE, base_dim=16)
    sd = V.make_weights(vcfg, seed=11, device="cuda")
    vae = AutoencoderKLWan(**vcfg).load_state_dict({k: v.clone() for k, v in sd.items()})

    class OracleVAE:
        dtype = torch.float32
        temperal_downsample = vcfg["temperal_downsample"]
        config = SimpleNamespace(**vcfg)

        def encode(self, x):
            m = V.encode_moments(x.double(), {k: v.double() for k, v in sd.items()}, vcfg, torch.float64).float()
            return SimpleNamespace(latent_dist=SimpleNamespace(mode=lambda: m[:, :16], sample=lambda generator=None: m[:, :16]))

        def decode(self, z, return_dict=True):
            v = V.decode(z.double(), {k: v.double() for k, v in sd.items()}, vcfg, torch.float64).float()
            return SimpleNamespace(sample=v) if return_dict else (v,)

    image = torch.rand(1, 3, 128, 192, generator=torch.Generator().manual_seed(3))
    frames = []
    for t, c, v in ((text, clip, vae), (hf_t, hf_c, OracleVAE())):
        pipe = WanImageToVideoPipeline(tokenizer=SyntheticTokenizer(vocab_size=4096), text_encoder=t, image_encoder=c,
                                       image_processor=SyntheticImageProcessor(), transformer=dit, vae=v,
                                       scheduler=UniPCMultistepScheduler(flow_shift=5.0)).to("cuda")
        out = pipe(image=image, prompt="a red bus turning a corner in the rain", negative_prompt="blurry", height=128, width=192,
                   num_frames=9, num_inference_steps=3, guidance_scale=5.0, max_sequence_length=32,
                   generator=torch.Generator(device="cuda").manual_seed(42), output_type="np", **alg)
        frames.append(np.asarray(out.frames))
    a, b = frames
    assert a.shape == b.shape == (1, 9, 128, 192, 3) and np.isfinite(a).all() and a.min() >= 0.0 and a.max() <= 1.0
    assert a.std() > 1e-3  # not a constant clip
    err = np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64))
    assert err < 2e-2, err


def test_gemm_tap_mode_is_a_stride1_convolution():
    """alg_gemm_bf16's implicit-convolution mode alone (bf16): 27 row-shifted reads of a zero-padded channels-last clip equal
    F.conv3d with causal zero padding; rows of the padded raster that are padding are don't-care."""
    import torch.nn.functional as F
    from alg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    T, H, W, Ci, Co = 3, 5, 7, 64, 96
    x = torch.randn(T, H, W, Ci, generator=g, device="cuda").bfloat16()
    w = (torch.randn(Co, Ci, 3, 3, 3, generator=g, device="cuda") * (27 * Ci) ** -0.5).bfloat16()
    pad = torch.zeros(T + 2, H + 2, W + 2, Ci, device="cuda", dtype=torch.bfloat16)
    pad[2:, 1:-1, 1:-1] = x
    a = pad.view(-1, Ci)
    wt = w.permute(0, 2, 3, 4, 1).reshape(Co, 27 * Ci).contiguous()
    plane, row = (H + 2) * (W + 2), W + 2
    offs = [(it - 2) * plane + (ih - 1) * row + (iw - 1) for it in range(3) for ih in range(3) for iw in range(3)]
    out = ops.gemm(a, wt, None, out_dtype=torch.float32, a_tap_kblocks=Ci // 64, a_tap_offsets=offs)
    got = out.view(T + 2, H + 2, W + 2, Co)[2:, 1:-1, 1:-1]
    ref = F.conv3d(F.pad(x.float().permute(3, 0, 1, 2)[None], (1, 1, 1, 1, 2, 0)), w.float())[0].permute(1, 2, 3, 0)
    assert rel_l2(got, ref) < 1e-5, rel_l2(got, ref)


def test_implicit_and_patch_matrix_paths_agree():
    """ALG_VAE_IMPLICIT=0 (gather + split patch matrix) and the default implicit convolutions give the same clip / moments."""
    vae, sd, ocfg = _pair(dict(TINY, z_dim=16), seed=5)
    z = torch.randn(1, 16, 3, 4, 6, generator=torch.Generator(device="cuda").manual_seed(7), device="cuda")
    x = torch.rand(1, 3, 5, 32, 48, generator=torch.Generator(device="cuda").manual_seed(8), device="cuda") * 2 - 1
    assert vae.implicit
    a, ma = vae.decode(z).sample, vae.encode(x).latent_dist.parameters
    vae.implicit = False
    b, mb = vae.decode(z).sample, vae.encode(x).latent_dist.parameters
    assert rel_l2(a, b) < 2e-5 and rel_l2(ma, mb) < 2e-5, (rel_l2(a, b), rel_l2(ma, mb))


def test_full_width_small_clip():
    """The real Wan2.1 VAE widths (96 / 192 / 384 channels, z = 16: the 320-column operand rows, the 96-wide GEMM tiles) on a small
    clip, encode and decode, implicit path."""
    from oracle import wan_vae_oracle as V
    vae, sd, ocfg = _pair(dict(V.WAN21_VAE), seed=9)
    sd64 = {k: v.double() for k, v in sd.items()}
    z = torch.randn(1, 16, 3, 6, 10, generator=torch.Generator(device="cuda").manual_seed(1), device="cuda")
    out = vae.decode(z).sample
    ref = V.decode(z.double(), sd64, ocfg, torch.float64)
    assert out.shape == ref.shape == (1, 3, 9, 48, 80) and rel_l2(out, ref) < 1e-4, rel_l2(out, ref)
    x = torch.rand(1, 3, 5, 48, 80, generator=torch.Generator(device="cuda").manual_seed(2), device="cuda") * 2 - 1
    m = vae.encode(x).latent_dist.parameters
    mref = V.encode_moments(x.double(), sd64, ocfg, torch.float64)
    assert rel_l2(m, mref) < 1e-4, rel_l2(m, mref)
