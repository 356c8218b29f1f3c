"""PyTorch (CPU) restatement of the three schedulers the reference drives.

TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED: the arithmetic lives in the third-party dependency
``diffusers @ git+https://github.com/huggingface/diffusers.git@be2fb77dc164083bf8f033874b066c96bc6752b8``
(/root/reference/requirements.txt:13), which is neither vendored under
/root/reference nor installed here.  What follows restates that library's
published algorithm (``scheduling_unipc_multistep.py``,
``scheduling_ddim_cogvideox.py``, ``scheduling_flow_match_euler_discrete.py``)
op by op -- including PyTorch's type-promotion of 0-dim tensors, which is what
decides where bf16 roundings happen -- anchored on the reference's call sites:

  * Wan:  ``UniPCMultistepScheduler.from_config(..., flow_shift=...)`` run.py:63;
          ``set_timesteps`` wan:815; ``step(noise_pred, t, latents)`` wan:927
  * Cog:  ``scheduler.step(noise_pred, t, latents, **extra)`` cog:1112; result ``.to(prompt_embeds.dtype)`` cog:1123
  * Hy:   ``FlowMatchEulerDiscreteScheduler.from_config(flow_shift, invert_sigmas)`` run.py:82-86;
          ``sigmas = linspace(1, 0, N+1)[:-1]`` hy:1111; ``step(noise[:, :, 1:], t, latents[:, :, 1:])`` hy:1265

CUDA-scalar convention.  The reference runs on CUDA, where ATen treats a 0-dim CPU
tensor (or Python number) operand of a CUDA tensor op as a *scalar at opmath (fp32)
precision*: ``sigma_t * bf16_tensor`` is ``bf16(fp32(sigma_t) * fp32(x))``, and
``tensor / scalar`` multiplies by the fp32 reciprocal (``div_true_kernel_cuda``).  Plain
PyTorch ops on CPU tensors would instead round the scalar to the tensor dtype first, so
this CPU restatement spells the CUDA behaviour out (``_smul`` / ``_div_scalar``); on CUDA
tensors the same helpers fall through to the plain ops, which is how the ``-m gpu`` tests
check that the spelling-out is faithful (measured: bit-identical).
"""
from __future__ import annotations

import math

import numpy as np
import torch


def _f32(s) -> torch.Tensor:
    return torch.as_tensor(s).detach().to("cpu", torch.float32)


def _smul(s, x: torch.Tensor) -> torch.Tensor:
    """``scalar * tensor`` the way ATen's CUDA kernels compute it (scalar enters at fp32, result in x.dtype)."""
    if x.is_cuda:
        return s * x
    return (x.to(torch.float32) * _f32(s)).to(x.dtype)


def _div_scalar(x: torch.Tensor, s) -> torch.Tensor:
    if x.is_cuda:
        return x / s
    return x * (torch.tensor(1.0, dtype=torch.float32) / _f32(s))


# ----------------------------------------------------------------------------
# UniPC (bh2, predict_x0, flow sigmas) as configured for Wan
# ----------------------------------------------------------------------------
class UniPCOracle:
    def __init__(self, num_train_timesteps=1000, solver_order=2, flow_shift=5.0, lower_order_final=True):
        self.num_train_timesteps = num_train_timesteps
        self.solver_order = solver_order
        self.flow_shift = flow_shift
        self.lower_order_final = lower_order_final
        self.order = 1

    def set_timesteps(self, n):
        alphas = np.linspace(1, 1 / self.num_train_timesteps, n + 1)
        sigmas = 1.0 - alphas
        sigmas = np.flip(self.flow_shift * sigmas / (1 + (self.flow_shift - 1) * sigmas))[:-1].copy()
        timesteps = (sigmas * self.num_train_timesteps).copy()
        sigmas = np.concatenate([sigmas, [0.0]]).astype(np.float32)  # final_sigmas_type == "zero"
        self.sigmas = torch.from_numpy(sigmas)
        self.timesteps = torch.from_numpy(timesteps).to(torch.int64)
        self.num_inference_steps = n
        self.model_outputs = [None] * self.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.step_index = 0
        self.this_order = None

    @staticmethod
    def _alpha_sigma(sigma):
        return 1 - sigma, sigma

    def _bh(self, sigma_t, sigma_s0, lambdas_prev, order):
        """Scalars shared by predictor/corrector: h_phi_1, B_h, rks, R, b."""
        alpha_t, sigma_t = self._alpha_sigma(sigma_t)
        alpha_s0, sigma_s0 = self._alpha_sigma(sigma_s0)
        lambda_t = torch.log(alpha_t) - torch.log(sigma_t)
        lambda_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = lambda_t - lambda_s0
        rks = []
        for lam in lambdas_prev[: order - 1]:
            rks.append((lam - lambda_s0) / h)
        rk_list = list(rks)
        rks.append(1.0)
        rks = torch.tensor(rks)
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        factorial_i = 1
        B_h = torch.expm1(hh)
        R, b = [], []
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * factorial_i / B_h)
            factorial_i *= i + 1
            h_phi_k = h_phi_k / hh - 1 / factorial_i
        R = torch.stack(R)
        b = torch.tensor(b)
        return alpha_t, sigma_t, sigma_s0, h_phi_1, B_h, rk_list, R, b

    def _lambda(self, idx):
        a, s = self._alpha_sigma(self.sigmas[idx])
        return torch.log(a) - torch.log(s)

    def step(self, model_output: torch.Tensor, sample: torch.Tensor) -> torch.Tensor:
        i = self.step_index
        use_corrector = i > 0 and self.last_sample is not None
        # convert_model_output (flow_prediction, predict_x0): 0-dim fp32 * bf16 tensor -> bf16
        sigma_t = self.sigmas[i]
        x0_pred = sample - _smul(sigma_t, model_output)
        if use_corrector:
            sample = self._uni_c(x0_pred, self.last_sample, sample, self.this_order)
        for k in range(self.solver_order - 1):
            self.model_outputs[k] = self.model_outputs[k + 1]
        self.model_outputs[-1] = x0_pred
        if self.lower_order_final:
            this_order = min(self.solver_order, len(self.timesteps) - i)
        else:
            this_order = self.solver_order
        self.this_order = min(this_order, self.lower_order_nums + 1)
        self.last_sample = sample
        prev = self._uni_p(sample, self.this_order)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self.step_index += 1
        return prev

    def _uni_p(self, x, order):
        i = self.step_index
        m0 = self.model_outputs[-1]
        lambdas_prev = [self._lambda(i - k) for k in range(1, order)]
        alpha_t, sigma_t, sigma_s0, h_phi_1, B_h, rks, R, b = self._bh(self.sigmas[i + 1], self.sigmas[i], lambdas_prev, order)
        D1s = [_div_scalar(self.model_outputs[-(k + 1)] - m0, rks[k - 1]) for k in range(1, order)]
        x_t_ = _smul(sigma_t / sigma_s0, x) - _smul(alpha_t * h_phi_1, m0)
        if D1s:
            assert order == 2
            rhos_p = torch.tensor([0.5], dtype=x.dtype)
            pred_res = _smul(rhos_p[0], D1s[0])
        else:
            pred_res = torch.zeros((), dtype=x.dtype, device=x.device)
        x_t = x_t_ - _smul(alpha_t * B_h, pred_res)
        return x_t.to(x.dtype)

    def _uni_c(self, model_t, x, this_sample, order):
        i = self.step_index
        m0 = self.model_outputs[-1]
        lambdas_prev = [self._lambda(i - (k + 1)) for k in range(1, order)]
        alpha_t, sigma_t, sigma_s0, h_phi_1, B_h, rks, R, b = self._bh(self.sigmas[i], self.sigmas[i - 1], lambdas_prev, order)
        D1s = [_div_scalar(self.model_outputs[-(k + 1)] - m0, rks[k - 1]) for k in range(1, order)]
        if order == 1:
            rhos_c = torch.tensor([0.5], dtype=x.dtype)
        else:
            rhos_c = torch.linalg.solve(R, b).to(x.dtype)
        x_t_ = _smul(sigma_t / sigma_s0, x) - _smul(alpha_t * h_phi_1, m0)
        if D1s:
            corr_res = _smul(rhos_c[0], D1s[0])
            for k in range(1, len(D1s)):
                corr_res = corr_res + _smul(rhos_c[k], D1s[k])
        else:
            corr_res = 0
        D1_t = model_t - m0
        x_t = x_t_ - _smul(alpha_t * B_h, corr_res + _smul(rhos_c[-1], D1_t))
        return x_t.to(x.dtype)


# ----------------------------------------------------------------------------
# CogVideoX DDIM (v-prediction, trailing spacing, zero-SNR rescale)
# ----------------------------------------------------------------------------
class CogDDIMOracle:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.0120, snr_shift_scale=1.0,
                 rescale_betas_zero_snr=True, set_alpha_to_one=True):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64) ** 2
        alphas = 1.0 - betas
        ac = torch.cumprod(alphas, dim=0)
        ac = ac / (snr_shift_scale + (1 - snr_shift_scale) * ac)
        if rescale_betas_zero_snr:
            s = ac.sqrt()
            s0, sT = s[0].clone(), s[-1].clone()
            s = (s - sT) * (s0 / (s0 - sT))
            ac = s ** 2
        self.alphas_cumprod = ac
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else ac[0]
        self.num_train_timesteps = num_train_timesteps
        self.order = 1

    def set_timesteps(self, n):
        self.num_inference_steps = n
        ts = np.round(np.arange(self.num_train_timesteps, 0, -self.num_train_timesteps / n)).astype(np.int64) - 1
        self.timesteps = torch.from_numpy(ts)

    def coeffs(self, t: int):
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        a = ((1 - a_prev) / b_t) ** 0.5
        b = a_prev ** 0.5 - a_t ** 0.5 * a
        return a_t, b_t, a, b

    def step(self, model_output: torch.Tensor, t: int, sample: torch.Tensor) -> torch.Tensor:
        a_t, b_t, a, b = self.coeffs(int(t))
        # 0-dim fp64 tensors do not promote dimensioned tensors: each product stays in its tensor's dtype
        pred_x0 = _smul(a_t ** 0.5, sample) - _smul(b_t ** 0.5, model_output)
        return _smul(a, sample) + _smul(b, pred_x0)


class CogDPMOracle(CogDDIMOracle):
    """``CogVideoXDPMScheduler.step`` (scheduling_dpm_cogvideox.py, v-prediction), restated op by op.  ``randn`` is the
    callable that draws the N(0, 1) tensors (the caller owns the generator): one draw for the first-order update and a
    second one when the second-order update replaces it."""

    def dpm_coeffs(self, t: int, t_back):
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        a_back = self.alphas_cumprod[t_back] if t_back is not None else None
        lamb = ((a_t / (1 - a_t)) ** 0.5).log()
        lamb_next = ((a_prev / (1 - a_prev)) ** 0.5).log()
        h = lamb_next - lamb
        mult = [((1 - a_prev) / (1 - a_t)) ** 0.5 * (-h).exp(), (-2 * h).expm1() * a_prev ** 0.5]
        if a_back is not None:
            r = (lamb - ((a_back / (1 - a_back)) ** 0.5).log()) / h
            mult += [1 + 1 / (2 * r), 1 / (2 * r)]
        mult_noise = (1 - a_prev) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5
        return a_t, prev_t, mult, mult_noise

    def step(self, model_output, old_pred, t: int, t_back, sample, randn):
        a_t, prev_t, mult, mult_noise = self.dpm_coeffs(int(t), None if t_back is None else int(t_back))
        pred = _smul(a_t ** 0.5, sample) - _smul((1 - a_t) ** 0.5, model_output)
        prev = _smul(mult[0], sample) - _smul(mult[1], pred) + _smul(mult_noise, randn())
        if old_pred is None or prev_t < 0:
            return prev, pred
        d = _smul(mult[2], pred) - _smul(mult[3], old_pred)
        return _smul(mult[0], sample) - _smul(mult[1], d) + _smul(mult_noise, randn()), pred


# ----------------------------------------------------------------------------
# FlowMatchEuler as configured for HunyuanVideo
# ----------------------------------------------------------------------------
class FlowEulerOracle:
    def __init__(self, num_train_timesteps=1000, shift=7.0, invert_sigmas=False):
        self.num_train_timesteps = num_train_timesteps
        self.shift = shift
        self.invert_sigmas = invert_sigmas
        self.order = 1

    def set_timesteps(self, n, sigmas=None):
        if sigmas is None:  # upstream spaces between the SHIFTED training schedule's end points (sigma_max, sigma_min)
            tr = (np.linspace(1, self.num_train_timesteps, self.num_train_timesteps, dtype=np.float32)[::-1] / self.num_train_timesteps)
            tr = self.shift * tr / (1 + (self.shift - 1) * tr)
            ts = np.linspace(float(tr[0]) * self.num_train_timesteps, float(tr[-1]) * self.num_train_timesteps, n)
            sigmas = ts / self.num_train_timesteps
        sigmas = np.array(sigmas).astype(np.float32)
        sigmas = self.shift * sigmas / (1 + (self.shift - 1) * sigmas)
        sigmas = torch.from_numpy(sigmas).to(torch.float32)
        timesteps = sigmas * self.num_train_timesteps
        if self.invert_sigmas:
            sigmas = 1.0 - sigmas
            timesteps = sigmas * self.num_train_timesteps
            sigmas = torch.cat([sigmas, torch.ones(1)])
        else:
            sigmas = torch.cat([sigmas, torch.zeros(1)])
        self.sigmas = sigmas
        self.timesteps = timesteps
        self.step_index = 0

    def step(self, model_output: torch.Tensor, sample: torch.Tensor) -> torch.Tensor:
        sample = sample.to(torch.float32)
        dt = self.sigmas[self.step_index + 1] - self.sigmas[self.step_index]
        if model_output.is_cuda:
            prev = sample + dt.to(model_output.device) * model_output
        else:
            # upstream keeps `sigmas` on the device, so dt is a 0-dim CUDA tensor: ATen casts it to the result dtype
            # (that of model_output) before the multiply -- unlike the CPU-scalar path of `_smul`
            dt_r = dt.to(model_output.dtype).to(torch.float32)
            prev = sample + (model_output.to(torch.float32) * dt_r).to(model_output.dtype)
        self.step_index += 1
        return prev.to(model_output.dtype)


# ----------------------------------------------------------------------------
# CFG combine (first-party: wan:919-924, cog:1091-1102, hy:1254-1261)
# ----------------------------------------------------------------------------
def cfg_combine(noise_pred: torch.Tensor, guidance_scale: float, fp32: bool = False) -> torch.Tensor:
    """3 chunks: u0 + w (t - u); 2 chunks: u + w (t - u).  Wan/Hy keep bf16 (three roundings); Cog ``.float()`` first."""
    if fp32:
        noise_pred = noise_pred.float()
    if noise_pred.shape[0] == 3:
        u0, u, t = noise_pred.chunk(3)
        return u0 + _smul(guidance_scale, t - u)
    u, t = noise_pred.chunk(2)
    return u + _smul(guidance_scale, t - u)
