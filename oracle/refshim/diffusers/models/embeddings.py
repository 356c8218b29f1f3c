"""``get_3d_rotary_pos_embed`` as cog:563-582 calls it (CogVideoX 1.0 branch: crops_coords + grid_size + temporal_size)."""
import torch


def _rope_1d(dim, pos, theta=10000.0):
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    f = torch.outer(pos, freqs)
    return f.cos().repeat_interleave(2, dim=1).float(), f.sin().repeat_interleave(2, dim=1).float()


def get_3d_rotary_pos_embed(embed_dim, crops_coords, grid_size, temporal_size, theta=10000, use_real=True,
                            grid_type="linspace", max_size=None, device=None):
    if grid_type != "linspace" or not use_real:
        raise NotImplementedError("refshim: only the CogVideoX 1.0 (linspace) tables are needed")
    start, stop = crops_coords
    gh_n, gw_n = grid_size
    gh = torch.linspace(start[0], stop[0] * (gh_n - 1) / gh_n, gh_n, dtype=torch.float32)
    gw = torch.linspace(start[1], stop[1] * (gw_n - 1) / gw_n, gw_n, dtype=torch.float32)
    gt = torch.arange(temporal_size, dtype=torch.float32)
    d_t, d_h, d_w = embed_dim // 4, embed_dim // 8 * 3, embed_dim // 8 * 3
    out = []
    for k in range(2):
        t = _rope_1d(d_t, gt, theta)[k][:, None, None, :].expand(-1, gh_n, gw_n, -1)
        h = _rope_1d(d_h, gh, theta)[k][None, :, None, :].expand(temporal_size, -1, gw_n, -1)
        w = _rope_1d(d_w, gw, theta)[k][None, None, :, :].expand(temporal_size, gh_n, -1, -1)
        out.append(torch.cat([t, h, w], dim=-1).reshape(temporal_size * gh_n * gw_n, -1).contiguous().to(device))
    return out[0], out[1]
