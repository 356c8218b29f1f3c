"""Fused CFG + scheduler.step kernels: bit-exact against the oracle (CPU, CUDA-scalar semantics spelled out) AND
against the plain eager op sequence on the GPU (the semantics the reference actually runs with)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
SHAPE = (1, 16, 5, 12, 20)


def _unipc(device, n_steps, n_pass, guidance, noise_dtype=torch.bfloat16):
    from alg_b200 import schedulers as S
    from oracle import sched_oracle as O
    torch.manual_seed(n_steps * 10 + n_pass)
    s = S.UniPCMultistepScheduler(flow_shift=5.0)
    s.set_timesteps(n_steps, device="cuda")
    o = O.UniPCOracle(flow_shift=5.0)
    o.set_timesteps(n_steps)
    x = torch.randn(SHAPE)
    xo, xg = x.to(device), x.cuda()
    for i in range(n_steps):
        npred = torch.randn((n_pass,) + SHAPE[1:]).to(noise_dtype)
        noise = O.cfg_combine(npred.to(device), guidance) if n_pass > 1 else npred.to(device)
        xo = o.step(noise, xo)
        xg = s.step_cfg(npred.cuda(), guidance, xg)
        assert torch.equal(xg.cpu(), xo.cpu()), f"step {i}: max diff {(xg.cpu() - xo.cpu()).abs().max()}"


@pytest.mark.parametrize("device", ["cpu", "cuda"])
@pytest.mark.parametrize("n_steps,n_pass,guidance", [(12, 3, 5.0), (12, 2, 5.0), (2, 3, 7.3), (1, 2, 5.0), (50, 3, 5.0), (6, 1, 1.0)])
def test_unipc_bit_exact(device, n_steps, n_pass, guidance):
    _unipc(device, n_steps, n_pass, guidance)


def test_unipc_fp32_noise():
    _unipc("cuda", 8, 3, 5.0, noise_dtype=torch.float32)


@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_ddim_bit_exact(device):
    from alg_b200 import schedulers as S
    from oracle import sched_oracle as O
    torch.manual_seed(1)
    d, od = S.CogVideoXDDIMScheduler(), O.CogDDIMOracle()
    d.set_timesteps(10, device="cuda")
    od.set_timesteps(10)
    x = torch.randn(SHAPE).bfloat16()
    xo, xg = x.to(device), x.cuda()
    for k, t in enumerate(od.timesteps):
        npass = 3 if k < 2 else 2
        npred = torch.randn((npass,) + SHAPE[1:]).bfloat16()
        noise = O.cfg_combine(npred.to(device), 6.0, fp32=True)
        xo = od.step(noise, int(t), xo).to(torch.bfloat16)
        xg = d.step_cfg(npred.cuda(), 6.0, int(t), xg)
        assert torch.equal(xg.cpu(), xo.cpu()), k


@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_euler_bit_exact(device):
    from alg_b200 import schedulers as S
    from oracle import sched_oracle as O
    torch.manual_seed(2)
    sig = torch.linspace(1, 0, 9)[:-1].numpy()
    e, oe = S.FlowMatchEulerDiscreteScheduler(shift=7.0), O.FlowEulerOracle(shift=7.0)
    e.set_timesteps(device="cuda", sigmas=sig)
    oe.set_timesteps(8, sigmas=sig)
    if device == "cuda":
        oe.sigmas = oe.sigmas.cuda()  # diffusers keeps FlowMatchEuler sigmas on the device
    x, first = torch.randn(SHAPE), torch.randn(1, 16, 1, 12, 20)
    xo, xg = x.to(device), x.cuda()
    for i in range(8):
        npass = 2 if i % 2 else 1
        npred = torch.randn((npass,) + SHAPE[1:]).bfloat16()
        noise = O.cfg_combine(npred.to(device), 6.0) if npass > 1 else npred.to(device)
        st = oe.step(noise[:, :, 1:], xo[:, :, 1:])
        xo = torch.cat([first.to(device), st], dim=2)
        xg = e.step_cfg_frames(npred.cuda(), 6.0, xg, first.cuda())
        assert xo.dtype == torch.float32 and torch.equal(xg.cpu(), xo.cpu()), i


def test_full_size_property_constant_velocity():
    """At the Wan config size (E = 2 096 640): a constant-velocity field is integrated back to x0 (size-independent)."""
    from alg_b200 import schedulers as S
    s = S.UniPCMultistepScheduler(flow_shift=5.0)
    s.set_timesteps(50, device="cuda")
    x0 = torch.randn(1, 16, 21, 60, 104, device="cuda")
    eps = torch.randn_like(x0)
    x = eps * float(s.sigmas[0]) + (1 - float(s.sigmas[0])) * x0
    for _ in range(50):
        x = s.step_cfg((eps - x0).unsqueeze(0).reshape(1, -1), 1.0, x)
    assert float((x - x0).abs().max()) < 1e-4


@pytest.mark.parametrize("n_pass", [1, 2, 3])
def test_cog_dpm_step_matches_oracle(n_pass):
    """cog:1113-1123: fp32 CFG + CogVideoXDPMScheduler.step (first step first-order, then second-order with the carried
    pred_original_sample, last step onto alpha = 1) against the oracle; both sides draw their noise from CUDA generators
    with the same seed, in the reference's order (one draw, plus a second one for the second-order update)."""
    from alg_b200 import schedulers as S
    from oracle import sched_oracle as O
    torch.manual_seed(3)
    steps, shape = 8, (1, 3, 16, 12, 20)
    e, oe = S.CogVideoXDPMScheduler(), O.CogDPMOracle()
    e.set_timesteps(steps, device="cuda")
    oe.set_timesteps(steps)
    g_e, g_o = (torch.Generator(device="cuda").manual_seed(9) for _ in range(2))
    x = torch.randn(shape, device="cuda").bfloat16()
    xo, old_e, old_o = x.clone(), None, None
    ts = e.timesteps.tolist()
    assert ts == oe.timesteps.tolist()
    for i, t in enumerate(ts):
        npred = torch.randn((n_pass,) + shape[1:], device="cuda").bfloat16()
        t_back = ts[i - 1] if i > 0 else None
        noise = O.cfg_combine(npred, 6.0, fp32=True) if n_pass > 1 else npred.float()
        xo, old_o = oe.step(noise, old_o, t, t_back, xo,
                            lambda: torch.randn(shape, generator=g_o, device="cuda", dtype=torch.bfloat16))
        xo = xo.to(torch.bfloat16)
        x, old_e = e.step_cfg(npred, 6.0, old_e, t, t_back, x, generator=g_e)
        assert old_e.dtype == torch.float32 and torch.equal(old_e, old_o), i
        assert x.dtype == torch.bfloat16 and torch.equal(x, xo), i
    assert torch.isfinite(x.float()).all()
