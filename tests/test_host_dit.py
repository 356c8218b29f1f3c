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
