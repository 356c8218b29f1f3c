"""Host-side position tables the CogVideoX / HunyuanVideo pipelines prepare once per call.

diffusers' ``models/embeddings.py`` (requirements.txt:13, @be2fb77) is not importable offline; these restate its
published functions as the reference calls them:

  * ``get_resize_crop_region_for_grid`` + ``get_3d_rotary_pos_embed``  <- cog:542-584 (CogVideoX 1.0 branch)
  * ``cogvideox_joint_pos_embedding``  <- CogVideoXPatchEmbed._get_positional_embeddings (3-D sincos, text part zero)
  * ``hunyuan_rotary_pos_embed``       <- HunyuanVideoRotaryPosEmbed (theta 256, axes 16/56/56) used at hy:1243

Pure host math on small tables (once per video, never per step); the tables are consumed by the CUDA kernels
(alg_head_norm_rope, the residual epilogue of the patch-embedding GEMM).
"""
from __future__ import annotations

from typing import Tuple

import torch


def get_resize_crop_region_for_grid(src, tgt_width, tgt_height):
    th, tw = tgt_height, tgt_width
    h, w = src
    r = h / w
    if r > (th / tw):
        resize_height = th
        resize_width = int(round(th / h * w))
    else:
        resize_width = tw
        resize_height = int(round(tw / w * h))
    crop_top = int(round((th - resize_height) / 2.0))
    crop_left = int(round((tw - resize_width) / 2.0))
    return (crop_top, crop_left), (crop_top + resize_height, crop_left + resize_width)


def get_1d_rotary_pos_embed(dim: int, pos: torch.Tensor, theta: float = 10000.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """use_real=True, repeat_interleave_real=True: cos, sin fp32 [len(pos), dim]."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32, device=pos.device)[: (dim // 2)] / dim))
    freqs = torch.outer(pos, freqs)
    return freqs.cos().repeat_interleave(2, dim=1).float(), freqs.sin().repeat_interleave(2, dim=1).float()


def _combine_thw(ft, fh, fw, T, H, W):
    ft = ft[:, None, None, :].expand(-1, H, W, -1)
    fh = fh[None, :, None, :].expand(T, -1, W, -1)
    fw = fw[None, None, :, :].expand(T, H, -1, -1)
    return torch.cat([ft, fh, fw], dim=-1).reshape(T * H * W, -1).contiguous()


def get_3d_rotary_pos_embed(embed_dim, crops_coords, grid_size, temporal_size, theta: int = 10000, device=None):
    """grid_type="linspace" (CogVideoX 1.0): cos, sin fp32 [temporal_size * gh * gw, embed_dim]."""
    start, stop = crops_coords
    gh_n, gw_n = grid_size
    grid_h = torch.linspace(start[0], stop[0] * (gh_n - 1) / gh_n, gh_n, device=device, dtype=torch.float32)
    grid_w = torch.linspace(start[1], stop[1] * (gw_n - 1) / gw_n, gw_n, device=device, dtype=torch.float32)
    grid_t = torch.arange(temporal_size, device=device, dtype=torch.float32)
    dim_t, dim_h, dim_w = embed_dim // 4, embed_dim // 8 * 3, embed_dim // 8 * 3
    t_cos, t_sin = get_1d_rotary_pos_embed(dim_t, grid_t, theta)
    h_cos, h_sin = get_1d_rotary_pos_embed(dim_h, grid_h, theta)
    w_cos, w_sin = get_1d_rotary_pos_embed(dim_w, grid_w, theta)
    return (_combine_thw(t_cos, h_cos, w_cos, temporal_size, gh_n, gw_n),
            _combine_thw(t_sin, h_sin, w_sin, temporal_size, gh_n, gw_n))


def _sincos_from_grid(embed_dim: int, pos: torch.Tensor) -> torch.Tensor:
    omega = torch.arange(embed_dim // 2, dtype=torch.float64) / (embed_dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = torch.outer(pos.reshape(-1).double(), omega)
    return torch.cat([torch.sin(out), torch.cos(out)], dim=1)


def cogvideox_joint_pos_embedding(dim, max_text_seq_length, grid_h, grid_w, latent_frames,
                                  spatial_interpolation_scale=1.875, temporal_interpolation_scale=1.0) -> torch.Tensor:
    """fp32 [1, max_text + T*h*w, dim]: zeros for the text tokens, [temporal quarter | spatial three quarters] sincos."""
    d_spatial, d_temporal = 3 * dim // 4, dim // 4
    gh = torch.arange(grid_h, dtype=torch.float32) / spatial_interpolation_scale
    gw = torch.arange(grid_w, dtype=torch.float32) / spatial_interpolation_scale
    mesh_w, mesh_h = torch.meshgrid(gw, gh, indexing="xy")
    spatial = torch.cat([_sincos_from_grid(d_spatial // 2, mesh_w), _sincos_from_grid(d_spatial // 2, mesh_h)], dim=1)
    temporal = _sincos_from_grid(d_temporal, torch.arange(latent_frames, dtype=torch.float32) / temporal_interpolation_scale)
    n_sp = grid_h * grid_w
    pos = torch.cat([temporal[:, None, :].expand(-1, n_sp, -1), spatial[None].expand(latent_frames, -1, -1)], dim=-1)
    joint = torch.zeros(1, max_text_seq_length + latent_frames * n_sp, dim, dtype=torch.float32)
    joint[0, max_text_seq_length:] = pos.reshape(-1, dim).float()
    return joint


def hunyuan_rotary_pos_embed(frames, grid_h, grid_w, rope_dim=(16, 56, 56), theta: float = 256.0, device=None):
    """HunyuanVideoRotaryPosEmbed.forward on a [T, h, w] token grid: cos, sin fp32 [T*h*w, sum(rope_dim)]."""
    axes = [torch.arange(0, n, device=device, dtype=torch.float32) for n in (frames, grid_h, grid_w)]
    grid = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=0)  # [3, T, h, w]
    cos, sin = [], []
    for i in range(3):
        c, s = get_1d_rotary_pos_embed(rope_dim[i], grid[i].reshape(-1), theta)
        cos.append(c)
        sin.append(s)
    return torch.cat(cos, dim=1).contiguous(), torch.cat(sin, dim=1).contiguous()
