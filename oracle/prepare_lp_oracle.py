"""Independent restatement of the three pipelines' ``prepare_lp`` and ``prepare_latents`` (SURVEY 8 rows a6 / boundary).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Nothing here imports ``alg_b200`` or the repo's ``pipeline_*`` / ``lp_utils``
modules: the filter is the reference's own two ATen / torchvision calls (``lp_utils.py:40-54``) spelled with
``F.interpolate`` / ``tvF.gaussian_blur``, and the surrounding tensor plumbing follows

  * Wan      ``prepare_lp`` wan:451-559, ``prepare_latents`` wan:372-449
  * CogVideoX ``prepare_lp`` cog:586-703, ``prepare_latents`` cog:351-431
  * Hunyuan  ``prepare_lp`` hy:650-792 (in-latent branch; the pixel branch is unreachable in the reference, quirk q9),
             ``prepare_latents`` hy:550-592

PINNED: ``oracle/gen_golden_loops.py`` runs the UNMODIFIED reference methods (imported from /root/reference through
``oracle/refshim``) on seeded inputs and stores what they return in ``tests/golden/loop_*.npz``;
``tests/test_oracle_loops_golden.py`` checks these restatements against those vectors.

``vae`` is any object with the diffusers surface (``encode(x).latent_dist.sample(generator) / .mode()``, ``config``); the
fixtures use ``oracle/stub_vae.ArithVAE``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
import torchvision.transforms.functional as tvF


# ----------------------------------------------------------------------------
# lp_utils.apply_low_pass_filter, lp_utils.py:8-60
# ----------------------------------------------------------------------------
def low_pass(x: torch.Tensor, kind: str, sigma: float, ksize, factor: float) -> torch.Tensor:
    if kind == "none" or (kind == "down_up" and factor == 1.0) or (kind == "gaussian_blur" and sigma == 0):
        return x  # same object (lp_utils.py:23-28)
    shape5 = x.shape if x.ndim == 5 else None
    if shape5 is not None:
        b, c, k, h, w = shape5
        x = x.view(b * k, c, h, w)  # no permute: planes are independent (lp_utils.py:31-35, quirk q6)
    h, w = x.shape[-2:]
    if kind == "gaussian_blur":
        k = max(int(ksize * h), 1) if isinstance(ksize, float) else int(ksize)  # a float is a fraction of H (quirk q8)
        k += 1 - k % 2
        x = tvF.gaussian_blur(x, kernel_size=[k, k], sigma=[sigma, sigma])
    elif kind == "down_up":
        small = (max(1, int(round(h * factor))), max(1, int(round(w * factor))))  # Python round: half to even (q14)
        x = F.interpolate(x, size=small, mode="bilinear", align_corners=False, antialias=True)
        x = F.interpolate(x, size=(h, w), mode="bilinear", align_corners=False, antialias=True)
    return x.view(shape5) if shape5 is not None else x


def _retrieve(enc, generator=None, mode="sample"):
    d = enc.latent_dist
    return d.sample(generator) if mode == "sample" else d.mode()


def _prepend_to_multiple(x, dim_size_of, multiple):
    """The "prepend frames to a multiple of the temporal patch" fix-up, exactly as written (tests dim 1; quirk q7)."""
    if multiple is None:
        return x
    rem = x.size(1) % multiple
    if rem:
        n = min(multiple - rem, x.shape[1])
        x = torch.cat([x[:, :n], x], dim=1)
    return x


# ----------------------------------------------------------------------------
# Wan
# ----------------------------------------------------------------------------
def _wan_norm(vae, device, dtype):
    z = vae.config.z_dim
    mean = torch.tensor(vae.config.latents_mean).view(1, z, 1, 1, 1).to(device, dtype)
    inv_std = 1.0 / torch.tensor(vae.config.latents_std).view(1, z, 1, 1, 1).to(device, dtype)
    return mean, inv_std


def _wan_mask(batch, num_frames, h_lat, w_lat, t_scale, device, keep_last):
    m = torch.ones(batch, 1, num_frames, h_lat, w_lat)
    m[:, :, 1:(num_frames - 1 if keep_last else num_frames)] = 0
    m = torch.cat([m[:, :, :1].repeat_interleave(t_scale, dim=2), m[:, :, 1:]], dim=2)
    return m.view(batch, -1, t_scale, h_lat, w_lat).transpose(1, 2).to(device)


def wan_prepare_latents(vae, image, batch_size, z_dim, height, width, num_frames, dtype, device, randn, generator=None,
                        latents=None, last_image=None, t_scale=4, s_scale=8):
    t_lat = (num_frames - 1) // t_scale + 1
    shape = (batch_size, z_dim, t_lat, height // s_scale, width // s_scale)
    latents = randn(shape, generator, device, dtype) if latents is None else latents.to(device=device, dtype=dtype)
    img = image.unsqueeze(2)
    gap = img.new_zeros(img.shape[0], img.shape[1], num_frames - (1 if last_image is None else 2), height, width)
    clip = torch.cat([img, gap] + ([] if last_image is None else [last_image.unsqueeze(2)]), dim=2).to(device=device, dtype=vae.dtype)
    mean, inv_std = _wan_norm(vae, latents.device, latents.dtype)
    if isinstance(generator, list):
        cond = torch.cat([_retrieve(vae.encode(clip), mode="argmax") for _ in generator])
    else:
        cond = _retrieve(vae.encode(clip), mode="argmax").repeat(batch_size, 1, 1, 1, 1)
    cond = (cond.to(dtype) - mean) * inv_std
    mask = _wan_mask(batch_size, num_frames, height // s_scale, width // s_scale, t_scale, cond.device, last_image is not None)
    return latents, torch.cat([mask, cond], dim=1)


def wan_prepare_lp(vae, patch_t, kind, sigma, ksize, factor, generator, num_frames, use_lp, in_latent, cond, image,
                   t_scale=4, s_scale=8):
    if not use_lp:
        return None
    if in_latent:
        lp = _prepend_to_multiple(low_pass(cond, kind, sigma, ksize, factor), 1, patch_t)
        return lp.to(cond.dtype)
    img_lp = low_pass(image, kind, sigma, ksize, factor)
    frame = img_lp.unsqueeze(2)
    b, _, height, width = image.shape
    clip = torch.cat([frame, frame.new_zeros(b, frame.shape[1], num_frames - 1, height, width)], dim=2)
    mean, inv_std = _wan_norm(vae, img_lp.device, img_lp.dtype)
    z = (vae.encode(clip).latent_dist.sample(generator=generator) - mean) * inv_std
    mask = _wan_mask(b, num_frames, height // s_scale, width // s_scale, t_scale, z.device, False)
    return torch.cat([mask, z], dim=1).to(cond.dtype)


# ----------------------------------------------------------------------------
# CogVideoX (latents are [B, F, C, H, W])
# ----------------------------------------------------------------------------
def _cog_scale(vae, z):
    s = vae.config.scaling_factor
    return (1 / s) * z if vae.config.invert_scale_latents else s * z


def cog_prepare_latents(vae, image, batch_size, channels, num_frames, height, width, dtype, device, randn, generator=None,
                        latents=None, patch_t=None, init_noise_sigma=1.0, t_scale=4, s_scale=8):
    f_lat = (num_frames - 1) // t_scale + 1
    shape = (batch_size, f_lat, channels, height // s_scale, width // s_scale)
    if patch_t is not None:
        shape = shape[:1] + (shape[1] + shape[1] % patch_t,) + shape[2:]
    img = image.unsqueeze(2)
    if isinstance(generator, list):
        enc = [_retrieve(vae.encode(img[i].unsqueeze(0)), generator[i]) for i in range(batch_size)]
    else:
        enc = [_retrieve(vae.encode(one.unsqueeze(0)), generator) for one in img]
    z = _cog_scale(vae, torch.cat(enc, dim=0).to(dtype).permute(0, 2, 1, 3, 4))
    pad = torch.zeros((batch_size, f_lat - 1, channels, height // s_scale, width // s_scale), device=device, dtype=dtype)
    z = torch.cat([z, pad], dim=1)
    if patch_t is not None:
        z = torch.cat([z[:, : z.size(1) % patch_t], z], dim=1)
    latents = randn(shape, generator, device, dtype) if latents is None else latents.to(device)
    return latents * init_noise_sigma, z


def cog_prepare_lp(vae, patch_t, kind, sigma, ksize, factor, generator, num_frames, use_lp, in_latent, image_latents, image,
                   t_scale=4):
    if not use_lp:
        return None
    if in_latent:
        lp = low_pass(image_latents.permute(0, 2, 1, 3, 4).contiguous(), kind, sigma, ksize, factor)
        lp = lp.permute(0, 2, 1, 3, 4).contiguous()
    else:
        img_lp = low_pass(image, kind, sigma, ksize, factor)
        z = _cog_scale(vae, vae.encode(img_lp.unsqueeze(2)).latent_dist.sample(generator=generator)).permute(0, 2, 1, 3, 4)
        want = (num_frames - 1) // t_scale + 1
        if want > z.shape[1]:
            b, have, c, h, w = z.shape
            z = torch.cat([z, torch.zeros((b, want - have, c, h, w), device=z.device, dtype=z.dtype)], dim=1)
        else:
            z = z[:, :want]
        lp = z
    return _prepend_to_multiple(lp, 1, patch_t).to(image_latents.dtype)


# ----------------------------------------------------------------------------
# HunyuanVideo (token_replace checkpoints)
# ----------------------------------------------------------------------------
def hunyuan_prepare_latents(vae, image, batch_size, channels, height, width, num_frames, dtype, device, randn,
                            generator=None, latents=None, image_condition_type="token_replace", i2v_stable=False,
                            t_scale=4, s_scale=8):
    t_lat = (num_frames - 1) // t_scale + 1
    shape = (batch_size, channels, t_lat, height // s_scale, width // s_scale)
    img = image.unsqueeze(2)
    if isinstance(generator, list):
        enc = [_retrieve(vae.encode(img[i].unsqueeze(0)), generator[i], "argmax") for i in range(batch_size)]
    else:
        enc = [_retrieve(vae.encode(one.unsqueeze(0)), generator, "argmax") for one in img]
    z = torch.cat(enc, dim=0).to(dtype) * vae.config.scaling_factor
    latents = randn(shape, generator, device, dtype) if latents is None else latents.to(device=device, dtype=dtype)
    if i2v_stable:
        z = z.repeat(1, 1, t_lat, 1, 1)
        t = torch.tensor([0.999]).to(device=device)
        latents = latents * t + z * (1 - t)
    if image_condition_type == "token_replace":
        z = z[:, :, :1]
    return latents, z


def hunyuan_prepare_lp(patch, kind, sigma, ksize, factor, use_lp, in_latent, image_latents):
    if not use_lp:
        return None
    if not in_latent:
        raise NotImplementedError("hy:698-770 passes a PIL image to the filter and reads Wan-VAE config fields: unreachable (q9)")
    return _prepend_to_multiple(low_pass(image_latents, kind, sigma, ksize, factor), 1, patch).to(image_latents.dtype)
