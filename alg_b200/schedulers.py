"""Schedulers of the three ALG pipelines with the CFG combine + ``step`` fused into one CUDA kernel.

They keep the ``diffusers`` scheduler surface the reference drives (``from_config``, ``config``,
``set_timesteps``, ``timesteps``, ``sigmas``, ``order``, ``step(model_output, timestep, sample)``) and add
``step_cfg(noise_pred_passes, guidance_scale, sample)`` which also folds in the reference's CFG arithmetic
(wan:919-927, cog:1091-1123, hy:1254-1270).  Only host-side SCALARS are computed here -- with the same torch
0-dim fp32 ops diffusers uses, so the coefficients are bit-identical -- the tensor update runs in
``libalg_b200.so``.  diffusers itself (requirements.txt:13, @be2fb77) is not importable offline; the algorithms are
restated from its published sources.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import List, Optional

import numpy as np
import torch

from . import _lib


class SchedulerOutput(tuple):
    @property
    def prev_sample(self):
        return self[0]


class _ConfigMixin:
    _defaults: dict = {}

    def __init__(self, **kwargs):
        cfg = dict(self._defaults)
        unknown = [k for k in kwargs if k not in cfg]
        for k in unknown:  # diffusers logs and ignores unexpected config keys (e.g. flow_shift for FlowMatchEuler)
            kwargs.pop(k)
        cfg.update(kwargs)
        self.config = SimpleNamespace(**cfg)
        self.ignored_config_keys = unknown

    @classmethod
    def from_config(cls, config, **kwargs):
        base = dict(vars(config)) if not isinstance(config, dict) else dict(config)
        base = {k: v for k, v in base.items() if k in cls._defaults}
        base.update(kwargs)
        return cls(**base)


def _as_passes(noise_pred: torch.Tensor, sample_numel: int):
    n = noise_pred.numel() // sample_numel
    if n * sample_numel != noise_pred.numel() or n not in (1, 2, 3):
        raise ValueError(f"noise_pred with {noise_pred.numel()} elements is not 1, 2 or 3 passes of {sample_numel}")
    return n


# =====================================================================================================
class UniPCMultistepScheduler(_ConfigMixin):
    """UniPC, bh2, predict-x0, flow sigmas -- the Wan configuration (run.py:63).  SURVEY Appendix B.1."""

    order = 1
    _defaults = dict(num_train_timesteps=1000, solver_order=2, prediction_type="flow_prediction", predict_x0=True,
                     solver_type="bh2", lower_order_final=True, use_flow_sigmas=True, flow_shift=1.0,
                     final_sigmas_type="zero", disable_corrector=())

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        c = self.config
        if not (c.prediction_type == "flow_prediction" and c.predict_x0 and c.solver_type == "bh2" and c.use_flow_sigmas):
            raise NotImplementedError("only the Wan configuration of UniPC (flow_prediction, predict_x0, bh2) is built")
        if c.solver_order not in (1, 2):
            raise NotImplementedError("solver_order must be 1 or 2")
        self.timesteps = None
        self.sigmas = None
        self._step_index = None

    # ---- host ------------------------------------------------------------------------------------
    def set_timesteps(self, num_inference_steps: int, device=None):
        c = self.config
        alphas = np.linspace(1, 1 / c.num_train_timesteps, num_inference_steps + 1)
        sigmas = 1.0 - alphas
        sigmas = np.flip(c.flow_shift * sigmas / (1 + (c.flow_shift - 1) * sigmas))[:-1].copy()
        timesteps = (sigmas * c.num_train_timesteps).copy()
        sigma_last = sigmas[-1] if c.final_sigmas_type == "sigma_min" else 0
        sigmas = np.concatenate([sigmas, [sigma_last]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas)  # stays on the host: only scalars are read from it
        self._sigmas_host = [float(x) for x in self.sigmas]
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self._timesteps_host = [int(t) for t in timesteps.astype(np.int64)]
        self.num_inference_steps = num_inference_steps
        self._step_index = 0
        self.lower_order_nums = 0
        self.this_order = None
        self._state = None  # device buffers: last_sample, m[2]
        self._cur = 0
        self._have_last = False
        # every step's scalar coefficients, computed ONCE here (the ~40 0-dim torch ops per step cost ~0.2 ms of host time,
        # 20x the fused kernel at the Wan config size): walk the order progression `step` will follow and memoise
        self._coef = {}
        lower, prev_order = 0, None
        for i in range(num_inference_steps):
            if i > 0:
                self._bh_at(i, i - 1, prev_order, True)
            order = min(c.solver_order, num_inference_steps - i) if c.lower_order_final else c.solver_order
            order = min(order, lower + 1)
            self._bh_at(i + 1, i, order, False)
            prev_order, lower = order, min(lower + 1, c.solver_order)

    @property
    def step_index(self):
        return self._step_index

    def _lambda(self, idx):
        s = self.sigmas[idx]
        return torch.log(1 - s) - torch.log(s)

    def _bh(self, sigma_t, sigma_s0, lambdas_prev, order, corrector: bool):
        alpha_t = 1 - sigma_t
        lambda_t = torch.log(alpha_t) - torch.log(sigma_t)
        lambda_s0 = torch.log(1 - sigma_s0) - torch.log(sigma_s0)
        h = lambda_t - lambda_s0
        rks = [(lam - lambda_s0) / h for lam in lambdas_prev[: order - 1]]
        rk_list = list(rks)
        rks = torch.tensor(rks + [1.0])
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
        R, b = torch.stack(R), torch.tensor(b)
        if corrector:
            rhos = torch.tensor([0.5]) if order == 1 else torch.linalg.solve(R, b).to(torch.float32)
        else:
            rhos = torch.tensor([0.5])  # order 2 uses the simplified rho_p; order 1 has no D1 term
        one = torch.tensor(1.0, dtype=torch.float32)
        return dict(ratio=float(sigma_t / sigma_s0), a=float(alpha_t * h_phi_1), b=float(alpha_t * B_h),
                    rk_inv=float(one / rk_list[0].to(torch.float32)) if rk_list else 0.0,
                    rho0=float(rhos[0]), rho_last=float(rhos[-1]))

    def _bh_at(self, i_t: int, i_s0: int, order: int, corrector: bool):
        """Coefficients of the update sigma[i_s0] -> sigma[i_t] (memoised by schedule position)."""
        key = (i_t, i_s0, order, corrector)
        k = self._coef.get(key)
        if k is None:
            lambdas_prev = [self._lambda(i_s0 - j) for j in range(1, order)]
            k = self._coef[key] = self._bh(self.sigmas[i_t], self.sigmas[i_s0], lambdas_prev, order, corrector)
        return k

    def _step_params(self, n_pass: int, guidance: float, cfg_fp32: bool) -> _lib.UniPCStep:
        i = self._step_index
        c = self.config
        p = _lib.UniPCStep()
        p.n_pass, p.cfg_fp32, p.guidance = n_pass, int(cfg_fp32), float(guidance)
        p.sigma_t = self._sigmas_host[i]
        use_corr = i > 0 and (i - 1) not in c.disable_corrector and self._have_last
        p.use_corrector = int(use_corr)
        if use_corr:
            oc = self.this_order
            k = self._bh_at(i, i - 1, oc, True)
            p.order_c, p.c_ratio, p.c_a, p.c_b = oc, k["ratio"], k["a"], k["b"]
            p.c_rk_inv, p.c_rho0, p.c_rho_last = k["rk_inv"], k["rho0"], k["rho_last"]
        else:
            p.order_c = 1
        if c.lower_order_final:
            this_order = min(c.solver_order, len(self._timesteps_host) - i)
        else:
            this_order = c.solver_order
        self.this_order = min(this_order, self.lower_order_nums + 1)
        op = self.this_order
        k = self._bh_at(i + 1, i, op, False)
        p.order_p, p.p_ratio, p.p_a, p.p_b, p.p_rk_inv, p.p_rho0 = op, k["ratio"], k["a"], k["b"], k["rk_inv"], 0.5
        return p

    # ---- device ----------------------------------------------------------------------------------
    def step_cfg(self, noise_pred: torch.Tensor, guidance_scale: float, sample: torch.Tensor, *, cfg_fp32=False,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """CFG over the 1/2/3 stacked passes of ``noise_pred`` + one UniPC step on ``sample`` (fp32)."""
        _lib.require_cuda(noise_pred, sample)
        if sample.dtype != torch.float32:
            raise TypeError("UniPC state is fp32 (the Wan pipeline keeps latents in float32, wan:825-837)")
        E = sample.numel()
        n_pass = _as_passes(noise_pred, E)
        noise_pred = noise_pred.contiguous()
        sample = sample.contiguous()
        if self._state is None or self._state[0].numel() != E or self._state[0].device != sample.device:
            self._state = [torch.zeros(E, device=sample.device, dtype=torch.float32) for _ in range(3)]
            self._cur, self._have_last = 0, False
        last, m = self._state[0], self._state[1:]
        p = self._step_params(n_pass, guidance_scale, cfg_fp32)
        if out is None:
            out = torch.empty_like(sample)
        m_prev0, m_prev1 = m[self._cur], m[1 - self._cur]
        with torch.cuda.device(sample.device):
            _lib.check(_lib.lib().alg_cfg_unipc_step(
                noise_pred.data_ptr(), _lib.dtype_code(noise_pred.dtype), sample.data_ptr(), out.data_ptr(),
                last.data_ptr(), m_prev0.data_ptr(), m_prev1.data_ptr(), E, C.byref(p), _lib.stream_ptr(sample.device)))
        self._cur = 1 - self._cur  # the buffer just written is now model_outputs[-1]
        self._have_last = True
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        return out

    def step(self, model_output, timestep=None, sample=None, return_dict: bool = True):
        prev = self.step_cfg(model_output, 1.0, sample)
        return SchedulerOutput((prev,))

    def scale_model_input(self, sample, *args, **kwargs):
        return sample


# =====================================================================================================
class CogVideoXDDIMScheduler(_ConfigMixin):
    """CogVideoX DDIM, v-prediction, trailing spacing, zero-terminal-SNR.  SURVEY Appendix B.2.

    ``_defaults`` are NOT diffusers' class defaults (epsilon / leading / snr_shift_scale 3.0 ...) but the values of the
    ``scheduler/scheduler_config.json`` THUDM/CogVideoX-5b-I2V ships (v_prediction, trailing, rescale_betas_zero_snr, snr_shift_scale
    1.0): the only configuration this class builds -- anything else raises below -- and what the synthetic pipelines must reproduce.
    Real snapshots carry every value explicitly (``from_config`` / ``checkpoint.scheduler_config``)."""

    order = 1
    init_noise_sigma = 1.0
    _defaults = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.0120, beta_schedule="scaled_linear",
                     clip_sample=False, set_alpha_to_one=True, steps_offset=0, prediction_type="v_prediction",
                     timestep_spacing="trailing", rescale_betas_zero_snr=True, snr_shift_scale=1.0)

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        c = self.config
        if c.prediction_type != "v_prediction" or c.beta_schedule != "scaled_linear" or c.timestep_spacing != "trailing":
            raise NotImplementedError("only the CogVideoX-5b-I2V DDIM configuration is built")
        betas = torch.linspace(c.beta_start ** 0.5, c.beta_end ** 0.5, c.num_train_timesteps, dtype=torch.float64) ** 2
        ac = torch.cumprod(1.0 - betas, dim=0)
        ac = ac / (c.snr_shift_scale + (1 - c.snr_shift_scale) * ac)
        if c.rescale_betas_zero_snr:
            s = ac.sqrt()
            s0, sT = s[0].clone(), s[-1].clone()
            s = (s - sT) * (s0 / (s0 - sT))
            ac = s ** 2
        self.alphas_cumprod = ac
        self.final_alpha_cumprod = torch.tensor(1.0) if c.set_alpha_to_one else ac[0]
        self.timesteps = None

    def set_timesteps(self, num_inference_steps: int, device=None):
        c = self.config
        self.num_inference_steps = num_inference_steps
        ts = np.round(np.arange(c.num_train_timesteps, 0, -c.num_train_timesteps / num_inference_steps)).astype(np.int64) - 1
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _coeffs(self, t: int):
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        a = ((1 - a_prev) / b_t) ** 0.5
        b = a_prev ** 0.5 - a_t ** 0.5 * a
        return float(a_t ** 0.5), float(b_t ** 0.5), float(a), float(b)

    def step_cfg(self, noise_pred, guidance_scale: float, timestep, sample, out=None):
        _lib.require_cuda(noise_pred, sample)
        E = sample.numel()
        n_pass = _as_passes(noise_pred, E)
        sa, sb, a, b = self._coeffs(int(timestep))
        noise_pred, sample = noise_pred.contiguous(), sample.contiguous()
        if out is None:
            out = torch.empty_like(sample)
        with torch.cuda.device(sample.device):
            _lib.check(_lib.lib().alg_cfg_ddim_step(
                noise_pred.data_ptr(), _lib.dtype_code(noise_pred.dtype), sample.data_ptr(), out.data_ptr(),
                _lib.dtype_code(sample.dtype), E, n_pass, float(guidance_scale), sa, sb, a, b,
                _lib.stream_ptr(sample.device)))
        return out

    def step(self, model_output, timestep, sample, eta: float = 0.0, generator=None, return_dict: bool = True, **kw):
        if eta != 0.0:
            raise NotImplementedError("eta > 0 (stochastic DDIM) is outside the hot path built here")
        return SchedulerOutput((self.step_cfg(model_output, 1.0, timestep, sample),))


# =====================================================================================================
class CogVideoXDPMScheduler(CogVideoXDDIMScheduler):
    """CogVideoX SDE DPM-Solver++ (2M) scheduler, v-prediction (cog:1113-1122).  Same betas / timesteps as the DDIM variant;
    ``step`` returns ``(prev_sample, pred_original_sample)`` and draws N(0, 1) noise on ``generator`` like the reference:
    once for the first-order update and once more when the second-order update replaces it (RNG stream position kept)."""

    def _dpm_coeffs(self, t: int, t_back: Optional[int]):
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        a_back = self.alphas_cumprod[t_back] if t_back is not None else None
        lamb = ((a_t / (1 - a_t)) ** 0.5).log()
        lamb_next = ((a_prev / (1 - a_prev)) ** 0.5).log()
        h = lamb_next - lamb
        m0 = ((1 - a_prev) / (1 - a_t)) ** 0.5 * (-h).exp()
        m1 = (-2 * h).expm1() * a_prev ** 0.5
        m_noise = (1 - a_prev) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5
        m2 = m3 = 0.0
        if a_back is not None:
            r = (lamb - ((a_back / (1 - a_back)) ** 0.5).log()) / h
            m2, m3 = 1 + 1 / (2 * r), 1 / (2 * r)
        return dict(sa=float(a_t ** 0.5), sb=float((1 - a_t) ** 0.5), m0=float(m0), m1=float(m1), m2=float(m2), m3=float(m3),
                    mn=float(m_noise), prev_t=prev_t)

    def step_cfg(self, noise_pred, guidance_scale: float, old_pred_original_sample, timestep, timestep_back, sample,
                 generator=None, out=None):
        """fp32 CFG over the stacked passes + one DPM step; returns (prev_sample [sample dtype], pred_original_sample fp32)."""
        from .pipeline_utils import randn_tensor

        _lib.require_cuda(noise_pred, sample, old_pred_original_sample)
        E = sample.numel()
        n_pass = _as_passes(noise_pred, E)
        k = self._dpm_coeffs(int(timestep), None if timestep_back is None else int(timestep_back))
        second = old_pred_original_sample is not None and k["prev_t"] >= 0
        rnd = randn_tensor(sample.shape, generator=generator, device=sample.device, dtype=sample.dtype)
        if second:  # the reference discards the first-order sample and draws again (cog scheduler `step`)
            rnd = randn_tensor(sample.shape, generator=generator, device=sample.device, dtype=sample.dtype)
        p = _lib.DpmStep()
        p.n_pass, p.second_order, p.guidance = n_pass, int(second), float(guidance_scale)
        p.sqrt_alpha_t, p.sqrt_beta_t = k["sa"], k["sb"]
        p.m0, p.m1, p.m2, p.m3, p.m_noise = k["m0"], k["m1"], k["m2"], k["m3"], k["mn"]
        noise_pred, sample = noise_pred.contiguous(), sample.contiguous()
        if out is None:
            out = torch.empty_like(sample)
        pred = torch.empty(sample.shape, device=sample.device, dtype=torch.float32)
        old = old_pred_original_sample.contiguous() if second else None
        if old is not None and old.dtype != torch.float32:
            raise TypeError("old_pred_original_sample is the fp32 tensor a previous step returned")
        with torch.cuda.device(sample.device):
            _lib.check(_lib.lib().alg_cfg_dpm_step(
                noise_pred.data_ptr(), _lib.dtype_code(noise_pred.dtype), sample.data_ptr(), out.data_ptr(),
                _lib.dtype_code(sample.dtype), None if old is None else old.data_ptr(), pred.data_ptr(), rnd.data_ptr(), E,
                C.byref(p), _lib.stream_ptr(sample.device)))
        return out, pred

    def step(self, model_output, old_pred_original_sample, timestep, timestep_back, sample, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None, variance_noise=None, return_dict: bool = False):
        return self.step_cfg(model_output, 1.0, old_pred_original_sample, timestep, timestep_back, sample, generator)


# =====================================================================================================
class FlowMatchEulerDiscreteScheduler(_ConfigMixin):
    """Flow-match Euler as configured for HunyuanVideo-I2V (run.py:82-86).  SURVEY Appendix B.3."""

    order = 1
    _defaults = dict(num_train_timesteps=1000, shift=7.0, use_dynamic_shifting=False, invert_sigmas=False)

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        if self.config.use_dynamic_shifting:
            raise NotImplementedError("dynamic shifting is not used by the HunyuanVideo-I2V configuration")
        c = self.config
        # diffusers keeps the SHIFTED training schedule's end points: set_timesteps(sigmas=None) spaces the timesteps between them
        train = (np.linspace(1, c.num_train_timesteps, c.num_train_timesteps, dtype=np.float32)[::-1] / c.num_train_timesteps).astype(np.float32)
        train = c.shift * train / (1 + (c.shift - 1) * train)
        self.sigma_max, self.sigma_min = float(train[0]), float(train[-1])
        self.timesteps = None
        self.sigmas = None
        self._step_index = None

    def set_timesteps(self, num_inference_steps: Optional[int] = None, device=None, sigmas: Optional[List[float]] = None):
        c = self.config
        if sigmas is None:  # linspace(sigma_to_t(sigma_max), sigma_to_t(sigma_min), n) / N -- then shifted (again) below, like diffusers
            ts = np.linspace(self.sigma_max * c.num_train_timesteps, self.sigma_min * c.num_train_timesteps, num_inference_steps)
            sigmas = ts / c.num_train_timesteps
        sigmas = np.array(sigmas).astype(np.float32)
        self.num_inference_steps = len(sigmas)
        sigmas = c.shift * sigmas / (1 + (c.shift - 1) * sigmas)
        sigmas = torch.from_numpy(sigmas).to(torch.float32)
        timesteps = sigmas * c.num_train_timesteps
        if c.invert_sigmas:
            sigmas = 1.0 - sigmas
            timesteps = sigmas * c.num_train_timesteps
            sigmas = torch.cat([sigmas, torch.ones(1)])
        else:
            sigmas = torch.cat([sigmas, torch.zeros(1)])
        self.sigmas = sigmas  # host
        self.timesteps = timesteps.to(device)
        self._step_index = 0

    @property
    def step_index(self):
        return self._step_index

    def step_cfg_frames(self, noise_pred, guidance_scale: float, sample, first_frame, out=None):
        """hy:1254-1270: (true) CFG + Euler on frames 1.. of [1, C, T, H, W] + re-prepend of ``first_frame``."""
        _lib.require_cuda(noise_pred, sample, first_frame)
        B, Cc, T, H, W = sample.shape
        if B != 1 or sample.dtype != torch.float32 or first_frame.dtype != torch.float32:
            raise NotImplementedError("HunyuanVideo loop: batch 1, fp32 latents container")
        E = sample.numel()
        n_pass = _as_passes(noise_pred, E)
        # diffusers keeps FlowMatchEuler sigmas ON THE DEVICE: `dt * model_output` is then a 0-dim CUDA tensor times a
        # bf16 tensor, and ATen casts the 0-dim operand to the result dtype (bf16) before multiplying
        dt = float((self.sigmas[self._step_index + 1] - self.sigmas[self._step_index]).to(noise_pred.dtype))
        noise_pred, sample, first_frame = noise_pred.contiguous(), sample.contiguous(), first_frame.contiguous()
        if out is None:
            out = torch.empty_like(sample)
        with torch.cuda.device(sample.device):
            _lib.check(_lib.lib().alg_cfg_euler_step(
                noise_pred.data_ptr(), _lib.dtype_code(noise_pred.dtype), sample.data_ptr(), out.data_ptr(),
                first_frame.data_ptr(), Cc, T, H * W, n_pass, float(guidance_scale), dt, _lib.stream_ptr(sample.device)))
        self._step_index += 1
        return out
