"""Real-checkpoint loader (alg_b200/checkpoint.py): a diffusers-layout snapshot written to disk (sharded safetensors +
index + config.json + scheduler_config.json) loads back into the CogVideoX / HunyuanVideo front-ends and through the
pipelines' ``from_pretrained`` -- CPU only (loading does not launch kernels); the Wan engine registers device pointers,
so its round trip is in the GPU suite."""
import json
import os

import pytest
import torch

COG_TINY = dict(num_attention_heads=2, attention_head_dim=64, time_embed_dim=64, text_embed_dim=64, num_layers=2,
                sample_width=12, sample_height=8, sample_frames=9, max_text_seq_length=16)


def _write_scheduler(snap, cfg):
    os.makedirs(os.path.join(snap, "scheduler"), exist_ok=True)
    with open(os.path.join(snap, "scheduler", "scheduler_config.json"), "w") as f:
        json.dump(dict(cfg, _class_name="X", _diffusers_version="0.34.0.dev0"), f)


def test_cog_snapshot_round_trip(tmp_path):
    from alg_b200 import checkpoint, cogvideox
    from oracle import cog_oracle as Co
    sd = Co.make_weights(Co.tiny_config(), dtype=torch.bfloat16, seed=5)
    snap = str(tmp_path / "snap")
    checkpoint.save_transformer(snap, dict(cogvideox.COGVIDEOX_5B_I2V, **COG_TINY), sd, "CogVideoXTransformer3DModel",
                                max_shard_bytes=200_000)  # forces several shards + an index file
    assert os.path.exists(os.path.join(snap, "transformer", checkpoint.INDEX_NAME))
    _write_scheduler(snap, dict(snr_shift_scale=2.0, timestep_spacing="trailing", some_future_key=1))
    model = cogvideox.CogVideoXTransformer3DModel.from_pretrained(snap, device="cpu")
    assert model.config.num_layers == 2 and model.config.max_text_seq_length == 16
    got = model.state_dict()
    assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    # hub-id form resolved inside an HF-cache layout
    cache = tmp_path / "cache" / "models--THUDM--CogVideoX-5b-I2V" / "snapshots"
    cache.mkdir(parents=True)
    os.symlink(snap, cache / "abc123")
    assert checkpoint.resolve_snapshot("THUDM/CogVideoX-5b-I2V", str(tmp_path / "cache")) == str(cache / "abc123")
    assert checkpoint.resolve_snapshot("THUDM/CogVideoX-5b-I2V", None) is None


def test_pipeline_from_pretrained_uses_snapshot(tmp_path):
    from alg_b200 import checkpoint, cogvideox
    from oracle import cog_oracle as Co
    from pipeline_cogvideox_image2video_lowpass import CogVideoXImageToVideoPipeline
    sd = Co.make_weights(Co.tiny_config(), dtype=torch.bfloat16, seed=6)
    snap = str(tmp_path / "snap")
    checkpoint.save_transformer(snap, dict(cogvideox.COGVIDEOX_5B_I2V, **COG_TINY), sd, "CogVideoXTransformer3DModel")
    _write_scheduler(snap, dict(snr_shift_scale=2.0))
    with pytest.raises(NotImplementedError, match="VAE"):
        CogVideoXImageToVideoPipeline.from_pretrained(snap, device="cpu")
    pipe = CogVideoXImageToVideoPipeline.from_pretrained(snap, device="cpu", allow_synthetic_aux=True)
    assert pipe.scheduler.config.snr_shift_scale == 2.0  # scheduler_config.json honoured, unknown keys ignored
    assert torch.equal(pipe.transformer.state_dict()["proj_out.weight"], sd["proj_out.weight"])
    with pytest.raises(FileNotFoundError, match="no local diffusers snapshot"):
        CogVideoXImageToVideoPipeline.from_pretrained("THUDM/CogVideoX-5b-I2V")


def test_hunyuan_snapshot_round_trip(tmp_path):
    from alg_b200 import checkpoint, hunyuan
    from oracle import hunyuan_oracle as Ho
    tiny = dict(num_attention_heads=2, attention_head_dim=128, num_layers=1, num_single_layers=1, num_refiner_layers=1,
                text_embed_dim=64, pooled_projection_dim=32)
    sd = Ho.make_weights(Ho.tiny_config(num_layers=1, num_single_layers=1), dtype=torch.bfloat16, seed=7)
    snap = str(tmp_path / "snap")
    checkpoint.save_transformer(snap, dict(hunyuan.HUNYUAN_VIDEO_I2V, **tiny), sd, "HunyuanVideoTransformer3DModel")
    model = hunyuan.HunyuanVideoTransformer3DModel.from_pretrained(snap, device="cpu")
    assert model.config.rope_axes_dim == (16, 56, 56)  # JSON lists come back as tuples
    assert all(torch.equal(model.state_dict()[k], sd[k]) for k in sd)


def test_cog_vae_encoder_snapshot(tmp_path):
    """A snapshot with a ``vae/`` folder: the pipeline builds the native VAE from it -- encoder (single-frame, the per-step
    path) and decoder -- and hands ``decode`` to the object passed as ``vae=`` when there is one (then only ``encoder.*`` is read)."""
    from safetensors.torch import save_file

    from alg_b200 import checkpoint, cogvideox, vae_cogvideox as V
    from alg_b200.pipeline_utils import SyntheticVideoVAE
    from oracle import cog_oracle as Co
    from pipeline_cogvideox_image2video_lowpass import CogVideoXImageToVideoPipeline
    snap = str(tmp_path / "snap")
    checkpoint.save_transformer(snap, dict(cogvideox.COGVIDEOX_5B_I2V, **COG_TINY),
                                Co.make_weights(Co.tiny_config(), dtype=torch.bfloat16, seed=6), "CogVideoXTransformer3DModel")
    _write_scheduler(snap, dict(snr_shift_scale=1.0))
    vcfg = dict(V.COGVIDEOX_5B_VAE, block_out_channels=[32, 64, 64, 64], layers_per_block=1, latent_channels=16)
    vsd = V.synthetic_state_dict(dict(vcfg, block_out_channels=tuple(vcfg["block_out_channels"])), seed=9, device="cpu")
    os.makedirs(os.path.join(snap, "vae"))
    with open(os.path.join(snap, "vae", "config.json"), "w") as f:
        json.dump(dict(vcfg, _class_name="AutoencoderKLCogVideoX", _diffusers_version="0.34.0.dev0", some_future_key=3), f)
    tcfg = dict(vcfg, block_out_channels=tuple(vcfg["block_out_channels"]))
    dsd = V.synthetic_decoder_state_dict(tcfg, seed=9, device="cpu")
    save_file(dict(vsd, **dsd), os.path.join(snap, "vae", checkpoint.WEIGHTS_NAME))

    full = V.AutoencoderKLCogVideoX.from_pretrained(snap, device="cpu")
    assert full.has_decoder and "decoder.up_blocks.0.resnets.0.norm1.yb.weight" in full._w  # conv_y | conv_b fused: [2 f, z]
    assert full._w["decoder.up_blocks.0.resnets.0.norm1.yb.weight"].shape == (128, 16)
    with pytest.raises(KeyError, match="missing decoder weights"):
        V.AutoencoderKLCogVideoX(**vcfg).load_state_dict(dict(vsd, **{"decoder.conv_in.conv.weight": dsd["decoder.conv_in.conv.weight"]}))
    enc = V.AutoencoderKLCogVideoX.from_pretrained(snap, device="cpu", decoder=SyntheticVideoVAE(z_dim=16))
    assert enc.config.block_out_channels == (32, 64, 64, 64) and enc.config.layers_per_block == 1
    w = vsd["encoder.down_blocks.1.resnets.0.conv1.conv.weight"]  # [64, 32, 3, 3, 3] -> GEMM operand [64, (kt, kh, kw, ci)]
    assert torch.equal(enc._w["encoder.down_blocks.1.resnets.0.conv1.conv.weight"], w.movedim(1, -1).reshape(64, -1))
    w_in = enc._w["encoder.conv_in.conv.weight"].view(32, 27, 8)  # 3 input channels zero-padded to 8
    assert torch.equal(w_in[..., :3], vsd["encoder.conv_in.conv.weight"].movedim(1, -1).reshape(32, 27, 3))
    assert not w_in[..., 3:].any()
    assert not any(k.startswith("decoder.") for k in enc._w)

    pipe = CogVideoXImageToVideoPipeline.from_pretrained(snap, device="cpu", allow_synthetic_aux=True)
    assert isinstance(pipe.vae, V.AutoencoderKLCogVideoX) and pipe.vae.has_decoder and pipe.vae.decoder is None
    assert pipe.vae_scale_factor_spatial == 8 and pipe.vae_scaling_factor_image == 0.7
    mine = SyntheticVideoVAE(z_dim=16)
    pipe2 = CogVideoXImageToVideoPipeline.from_pretrained(snap, device="cpu", vae=mine, allow_synthetic_aux=True)
    assert pipe2.vae.decoder is mine
    pipe3 = CogVideoXImageToVideoPipeline.from_pretrained(snap, device="cpu", vae=mine, native_vae_encoder=False, allow_synthetic_aux=True)
    assert pipe3.vae is mine
    with pytest.raises(KeyError, match="missing encoder weights"):
        V.AutoencoderKLCogVideoX(**dict(vcfg, layers_per_block=2)).load_state_dict(vsd)
