"""CPU checks of the host-side DiT sequencers: alg_b200/cogvideox.py and alg_b200/hunyuan.py only order C-ABI calls, so
with ``ops`` swapped for the eager emulation (oracle/ops_emulation.py, test infrastructure) they must reproduce the model
oracles -- weight naming, joint-buffer layout, text/video row splits, modulation chunk order, RoPE tables, unpatchify
order -- without a GPU.  The product path itself has no CPU mode: without the swap every op raises."""
import pytest
import torch

from conftest import rel_l2

COG_TINY = dict(num_attention_heads=2, attention_head_dim=64, time_embed_dim=64, text_embed_dim=64, num_layers=2,
                sample_width=12, sample_height=8, sample_frames=9, max_text_seq_length=16)


def test_cog_sequencer_matches_oracle(monkeypatch):
    from alg_b200 import cogvideox
    from oracle import cog_oracle as Co, ops_emulation as emu
    monkeypatch.setattr(cogvideox, "ops", emu)
    cfg = Co.tiny_config()
    sd = Co.make_weights(cfg, dtype=torch.bfloat16, seed=1)
    g = torch.Generator().manual_seed(0)
    Fr, H, W = 3, 8, 12
    x = torch.randn(3, Fr, 32, H, W, generator=g).bfloat16()
    text = torch.randn(3, 16, 64, generator=g).bfloat16()
    t = torch.tensor([999] * 3)
    rope = Co.rotary_tables(cfg, H // 2, W // 2, Fr)
    model = cogvideox.CogVideoXTransformer3DModel(**COG_TINY).load_state_dict(sd)
    out = model(x, text, t, image_rotary_emb=rope, return_dict=False)[0]
    ref = Co.forward(sd, cfg, x, text, t, rope)
    assert rel_l2(out, ref) < 2e-3, rel_l2(out, ref)
    # [neg, neg, pos]: passes that share a prompt tensor reuse its projection
    lat = [x[b, :, :16] for b in range(3)]
    img = [x[b, :, 16:] for b in range(3)]
    out2 = model.forward_passes(lat, img, [text[0], text[0], text[2]], 999, rope)
    ref2 = Co.forward(sd, cfg, x, torch.stack([text[0], text[0], text[2]]), t, rope)
    assert rel_l2(out2, ref2) < 2e-3
    # another frame count: the sincos positional table is rebuilt on the host like CogVideoXPatchEmbed does
    x5 = torch.randn(1, 5, 32, H, W, generator=g).bfloat16()
    rope5 = Co.rotary_tables(cfg, H // 2, W // 2, 5)
    out5 = model(x5, text[:1], torch.tensor([500]), image_rotary_emb=rope5, return_dict=False)[0]
    assert rel_l2(out5, Co.forward(sd, cfg, x5, text[:1], torch.tensor([500]), rope5)) < 2e-3


def test_cog_host_tables_match_oracle():
    from alg_b200 import embeddings
    from oracle import cog_oracle as Co
    cfg = Co.CogConfig()
    cos, sin = embeddings.get_3d_rotary_pos_embed(64, embeddings.get_resize_crop_region_for_grid((30, 45), 45, 30), (30, 45), 13)
    rc, rs = Co.rotary_tables(cfg, 30, 45, 13)
    assert cos.shape == (13 * 30 * 45, 64) and torch.equal(cos, rc) and torch.equal(sin, rs)
    tab = embeddings.cogvideox_joint_pos_embedding(3072, 226, 30, 45, 3)
    ref = Co.joint_pos_embedding(cfg, 60, 90, 3)
    assert tab.shape == ref.shape == (1, 226 + 3 * 1350, 3072)
    assert (tab - ref).abs().max() < 1e-6 and not tab[:, :226].any()


def test_ops_fail_loudly_without_cuda():
    from alg_b200 import ops
    x = torch.zeros(4, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.layer_norm(x, eps=1e-5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.head_norm_rope(x, 1, 64)


HY_TINY = dict(num_attention_heads=2, attention_head_dim=128, num_layers=2, num_single_layers=2, num_refiner_layers=1,
               text_embed_dim=64, pooled_projection_dim=32)


def test_hunyuan_sequencer_matches_oracle(monkeypatch):
    """Also checks that dropping the padded text tokens (the engine's form of the key-padding mask) is exact: the
    oracle keeps them and builds diffusers' boolean masks."""
    from alg_b200 import hunyuan
    from oracle import hunyuan_oracle as Ho, ops_emulation as emu
    monkeypatch.setattr(hunyuan, "ops", emu)
    cfg = Ho.tiny_config()
    sd = Ho.make_weights(cfg, dtype=torch.bfloat16, seed=2)
    g = torch.Generator().manual_seed(0)
    T, H, W = 3, 8, 16
    x = torch.randn(2, 16, T, H, W, generator=g)
    text = torch.randn(2, 12, 64, generator=g).bfloat16()
    mask = torch.zeros(2, 12)
    mask[0, :7] = 1
    mask[1, :10] = 1
    pooled = torch.randn(2, 32, generator=g).bfloat16()
    t = torch.tensor([900.0, 900.0]).bfloat16()
    gd = torch.tensor([6.0, 6.0]).bfloat16() * 1000.0
    assert float(gd[0]) == 6016.0  # hy:1115-1119: bf16(6.0) * 1000 rounds to 6016
    model = hunyuan.HunyuanVideoTransformer3DModel(**HY_TINY).load_state_dict(sd)
    out = model(x, t, text, mask, pooled, gd, return_dict=False)[0]
    ref = Ho.forward(sd, cfg, x.bfloat16(), t, text, mask, pooled, gd)
    ref32 = Ho.forward({k: v.float() for k, v in sd.items()}, cfg, x.bfloat16().float(), t.float(), text.float(), mask,
                       pooled.float(), gd.float())
    assert rel_l2(out, ref) < 2e-3, rel_l2(out, ref)
    assert rel_l2(out, ref32) < 1.2 * rel_l2(ref, ref32)
    # first-frame pointer override == cat([first, latents[:, :, 1:]], dim=2) (hy:1171, 1232)
    first = torch.randn(16, 1, H, W, generator=g)
    o2 = model.forward_pass(x[0], first, text[0, :7].contiguous(), pooled[0], float(t[0]), float(gd[0]))
    x2 = x[:1].clone()
    x2[0, :, 0:1] = first
    r2 = Ho.forward(sd, cfg, x2.bfloat16(), t[:1], text[:1], mask[:1], pooled[:1], gd[:1])
    assert rel_l2(o2, r2[0]) < 2e-3


def test_hunyuan_rope_table_matches_oracle():
    from alg_b200 import embeddings
    from oracle import hunyuan_oracle as Ho
    cos, sin = embeddings.hunyuan_rotary_pos_embed(5, 6, 10)
    rc, rs = Ho.rope_tables(Ho.HunyuanConfig(), 5, 6, 10)
    assert cos.shape == (300, 128) and torch.equal(cos, rc) and torch.equal(sin, rs)
