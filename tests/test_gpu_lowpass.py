"""CUDA low-pass kernels (through the C ABI) vs the reference fixtures, the numpy oracle and -- for 16-bit tensors,
which the CPU reference cannot run (quirk q16) -- vs the reference's own op (F.interpolate on the GPU)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _cases(path):
    g = np.load(path)
    return g, sorted({k[:-2] for k in g.files if k.endswith("_x")})


def test_down_up_matches_reference_fixtures(golden_dir):
    import lp_utils
    g, names = _cases(f"{golden_dir}/lp_down_up.npz")
    for n in names:
        x = torch.from_numpy(g[n + "_x"]).cuda()
        y = lp_utils.apply_low_pass_filter(x, "down_up", 0.0, 0.0, float(g[n + "_f"]))
        assert y.shape == x.shape and y.dtype == x.dtype
        assert rel_l2(y, torch.from_numpy(g[n + "_y"])) < 2e-6, n  # fp32 tolerance stated in SURVEY 8(c)


def test_gaussian_matches_reference_fixtures(golden_dir):
    import lp_utils
    g, names = _cases(f"{golden_dir}/lp_gaussian.npz")
    for n in names:
        dt = getattr(torch, str(g[n + "_dtype"]))
        k = g[n + "_k"]
        k = float(k) if k.dtype == np.float64 else int(k)
        x = torch.from_numpy(g[n + "_x"]).to(dt).cuda()
        y = lp_utils.apply_low_pass_filter(x, "gaussian_blur", float(g[n + "_sigma"]), k, 0.0)
        ref = torch.from_numpy(g[n + "_y"])
        d = y.float().cpu() - ref
        if dt == torch.float32:
            assert rel_l2(y, ref) < 2e-6, n
        else:  # bit-exact up to fp32 summation order inside conv2d: <= 1 bf16 ulp on < 0.1 % of the elements
            assert float((d != 0).float().mean()) < 1e-3, n
            assert float(d.abs().max()) <= float(ref.abs().max()) * 2 ** -7, n


@pytest.mark.parametrize("shape,f", [((1, 20, 21, 60, 104), 0.4), ((1, 16, 1, 90, 160), 0.625), ((3, 5, 33, 47), 0.3),
                                      ((1, 3, 480, 832), 0.25), ((2, 1, 7, 9), 0.5), ((1, 2, 64, 64), 1.7)])
def test_down_up_matches_torch_and_oracle_fp32(shape, f):
    import lp_utils
    from oracle import lp_oracle
    x = torch.randn(shape, device="cuda")
    y = lp_utils.apply_low_pass_filter(x, "down_up", 0.0, 0.0, f)
    H, W = shape[-2:]
    h1, w1 = max(1, int(round(H * f))), max(1, int(round(W * f)))
    v = x.reshape(-1, 1, H, W)
    r = F.interpolate(F.interpolate(v, size=(h1, w1), mode="bilinear", antialias=True), size=(H, W), mode="bilinear", antialias=True)
    assert rel_l2(y, r.view(shape)) < 4e-6  # two fp32 evaluation orders of the same taps (ATen itself is 1.8e-6 off fp64)
    if x.numel() < 1_000_000:
        o = lp_oracle.apply_low_pass_filter(x.cpu().numpy(), "down_up", 0.0, 0.0, f)
        assert rel_l2(y, torch.from_numpy(o)) < 2e-6


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape,f", [((1, 16, 13, 60, 90), 0.25), ((1, 4, 2, 60, 104), 0.4), ((2, 3, 31, 45), 0.6)])
def test_down_up_16bit_matches_reference_op_on_gpu(dtype, shape, f):
    """The shipped CogVideoX ALG config filters a bf16 latent on the GPU (cog:684-692): the kernel reproduces ATen's
    rounding points (taps in the tensor dtype, row pass and small image materialised in the tensor dtype)."""
    import lp_utils
    x = torch.randn(shape, device="cuda").to(dtype)
    y = lp_utils.apply_low_pass_filter(x, "down_up", 0.0, 0.0, f)
    H, W = shape[-2:]
    h1, w1 = max(1, int(round(H * f))), max(1, int(round(W * f)))
    v = x.reshape(-1, 1, H, W)
    r = F.interpolate(F.interpolate(v, size=(h1, w1), mode="bilinear", antialias=True), size=(H, W), mode="bilinear", antialias=True).view(shape)
    d = (y.float() - r.float())
    ulp = 2 ** -7 if dtype == torch.bfloat16 else 2 ** -10
    assert float((d != 0).float().mean()) < 2e-2, float((d != 0).float().mean())
    assert float(d.abs().max()) <= float(r.float().abs().max()) * ulp


def test_gaussian_matches_torchvision_on_gpu():
    import lp_utils
    import torchvision.transforms.functional as tvF
    for dt, shape, k, s in ((torch.bfloat16, (1, 3, 480, 720), 13, 15.0), (torch.float32, (1, 3, 480, 832), 13, 7.5),
                            (torch.float16, (2, 3, 65, 70), 9, 2.0), (torch.float32, (1, 2, 3, 40, 40), 5, 1.0)):
        x = torch.randn(shape, device="cuda").to(dt)
        y = lp_utils.apply_low_pass_filter(x, "gaussian_blur", s, k, 0.0)
        r = tvF.gaussian_blur(x.reshape(-1, shape[-3], shape[-2], shape[-1]), [k, k], [s, s]).view(shape)
        d = y.float() - r.float()
        if dt == torch.float32:
            assert rel_l2(y, r) < 2e-6
        else:
            assert float((d != 0).float().mean()) < 1e-2
            assert float(d.abs().max()) <= float(r.float().abs().max()) * (2 ** -7 if dt == torch.bfloat16 else 2 ** -10)


def test_domain_properties_at_config_size():
    import lp_utils
    # constants survive (Wan mask channels, quirk q5); linearity; planes are independent (quirk q6)
    c = torch.full((1, 4, 21, 60, 104), 1.0, device="cuda")
    assert float((lp_utils.apply_low_pass_filter(c, "down_up", 0, 0, 0.4) - 1).abs().max()) < 1e-6
    a, b = torch.randn(2, 1, 20, 21, 60, 104, device="cuda")
    fa, fb = (lp_utils.apply_low_pass_filter(t, "down_up", 0, 0, 0.4) for t in (a, b))
    fab = lp_utils.apply_low_pass_filter(2 * a - 3 * b, "down_up", 0, 0, 0.4)
    assert rel_l2(fab, 2 * fa - 3 * fb) < 1e-5
    one = lp_utils.apply_low_pass_filter(a[:, 3:4, 5:6].contiguous(), "down_up", 0, 0, 0.4)
    assert torch.equal(one, fa[:, 3:4, 5:6])
    # low-pass: a second application changes little compared with the first
    ffa = lp_utils.apply_low_pass_filter(fa, "down_up", 0, 0, 0.4)
    assert rel_l2(ffa, fa) < 0.5 * rel_l2(fa, a)


def test_edge_cases():
    import lp_utils
    x = torch.randn(1, 2, 8, 8, device="cuda")
    assert lp_utils.apply_low_pass_filter(x, "none", 1.0, 3, 0.5) is x
    assert lp_utils.apply_low_pass_filter(x, "down_up", 1.0, 3, 1.0) is x
    assert lp_utils.apply_low_pass_filter(x, "gaussian_blur", 0, 3, 0.5) is x
    e = torch.empty(0, 3, 8, 8, device="cuda")
    assert lp_utils.apply_low_pass_filter(e, "down_up", 0, 0, 0.5).shape == e.shape
    nc = torch.randn(1, 8, 2, 8, device="cuda").permute(0, 2, 1, 3)  # non-contiguous input
    y = lp_utils.apply_low_pass_filter(nc, "down_up", 0, 0, 0.5)
    assert rel_l2(y, lp_utils.apply_low_pass_filter(nc.contiguous(), "down_up", 0, 0, 0.5)) == 0
    with pytest.raises(RuntimeError, match="reflect padding"):
        lp_utils.apply_low_pass_filter(torch.randn(1, 1, 4, 4, device="cuda"), "gaussian_blur", 1.0, 9, 0.5)


@pytest.mark.parametrize("H,W,k,sigma", [(480, 832, 13, 15.0), (33, 70, 5, 1.3), (64, 64, 31, 7.0), (40, 19, 3, 0.8), (130, 200, 63, 20.0)])
def test_gaussian_fp32_separable_matches_torchvision(H, W, k, sigma):
    """fp32 tensors take the separable kernel (row pass + column pass): within 2e-6 of torchvision's dense [k, k] depthwise
    convolution with reflect padding, on ragged sizes, tile-edge windows (k % 4 = 1, 3), the largest kernel, several planes."""
    import torchvision.transforms.functional as tvF
    import lp_utils
    g = torch.Generator(device="cuda").manual_seed(H * W + k)
    x = torch.randn(2, 3, H, W, generator=g, device="cuda")
    out = lp_utils.apply_low_pass_filter(x, "gaussian_blur", sigma, k, 0.25)
    ref = tvF.gaussian_blur(x.double(), kernel_size=[k, k], sigma=[sigma, sigma])
    assert out.dtype == torch.float32 and out.shape == x.shape
    assert rel_l2(out, ref) < 2e-6, rel_l2(out, ref)
