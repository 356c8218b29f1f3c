"""Native conditioning encoders (alg_b200/encoders.py over the C ABI) against the REAL transformers modules with seeded
random weights (transformers is installed here, unlike diffusers, so this row's oracle is the upstream code itself).

Tolerances.  UMT5 / T5 run in bf16 (run.py loads the text encoder in the model dtype): like the DiTs, the engine must be no
further from an fp32 evaluation of the same bf16-rounded weights than transformers' own bf16 forward is (x1.5), and within
2e-2 of that bf16 forward.  CLIP-ViT runs in float32 (run.py:48): the bf16x3-split tensor-core linears are required within
1e-4 relative L2 of transformers' fp32 forward at every hidden state."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _umt5(cls_name, cfg_over, seed=0):
    import transformers
    from transformers import T5Config, UMT5Config
    torch.manual_seed(seed)
    base = dict(vocab_size=512, d_model=128, d_kv=32, d_ff=256, num_layers=2, num_heads=4, relative_attention_num_buckets=32,
                relative_attention_max_distance=128, feed_forward_proj="gated-gelu", dense_act_fn="gelu_new")
    base.update(cfg_over)
    cfg = (UMT5Config if cls_name == "UMT5EncoderModel" else T5Config)(**base)
    hf = getattr(transformers, cls_name)(cfg).eval()
    with torch.no_grad():  # non-trivial norms / bias tables (the default init is all-ones / tiny)
        for n, p in hf.named_parameters():
            if "layer_norm" in n:
                p.copy_(1 + 0.1 * torch.randn_like(p))
            elif "relative_attention_bias" in n:
                p.copy_(torch.randn_like(p))
            elif p.dim() == 2 and "shared" not in n and "embed_tokens" not in n:
                p.copy_(torch.randn_like(p) * p.shape[1] ** -0.5)
    return base, hf


def _ids(B, L, vocab, lens, seed=1):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(2, vocab, (B, L), generator=g)
    mask = torch.zeros(B, L, dtype=torch.long)
    for b, n in enumerate(lens):
        mask[b, :n] = 1
        ids[b, n:] = 0
    return ids.cuda(), mask.cuda()


@pytest.mark.parametrize("cls_name,over,L,lens", [
    ("UMT5EncoderModel", {}, 24, (24, 9)),
    ("UMT5EncoderModel", dict(num_heads=3, d_kv=64, d_model=192, d_ff=384, num_layers=3), 130, (130, 77)),
    ("T5EncoderModel", {}, 40, (17, 40)),
    ("UMT5EncoderModel", dict(d_model=4096, d_kv=64, num_heads=64, d_ff=10240, num_layers=2, vocab_size=1024), 512, (301, 512)),
])
def test_text_encoder_matches_transformers(cls_name, over, L, lens):
    from alg_b200 import encoders
    base, hf = _umt5(cls_name, over)
    hf16 = hf.to(torch.bfloat16).cuda()
    sd = {k: v.detach().clone() for k, v in hf16.state_dict().items()}
    mine = getattr(encoders, cls_name)(per_layer_relative_bias=cls_name == "UMT5EncoderModel", **{k: v for k, v in base.items()})
    mine.load_state_dict(sd)
    ids, mask = _ids(2, L, base["vocab_size"], lens)
    with torch.no_grad():
        ref16, ref16_nomask = hf16(ids, mask).last_hidden_state, hf16(ids).last_hidden_state
        hf32 = hf16.float()  # in place: same bf16-rounded weights, fp32 arithmetic
        ref32, ref32_nomask = hf32(ids, mask).last_hidden_state, hf32(ids).last_hidden_state
    out = mine(ids, mask).last_hidden_state
    assert out.shape == ref16.shape and out.dtype == torch.bfloat16
    for b, n in enumerate(lens):  # padded query rows are discarded by the pipeline (wan:214-217); compare the prompt's own rows
        e_mine, e_hf = rel_l2(out[b, :n], ref32[b, :n]), rel_l2(ref16[b, :n], ref32[b, :n])
        assert e_mine < max(1.5 * e_hf, 4e-3), (b, e_mine, e_hf)
        assert rel_l2(out[b, :n], ref16[b, :n]) < max(2e-2, e_hf)  # two bf16 evaluations are each e_hf away from fp32
    # no mask at all (cog:258 calls the encoder with ids only)
    out_nomask = mine(ids).last_hidden_state
    assert rel_l2(out_nomask, ref32_nomask) < max(1.5 * rel_l2(ref16_nomask, ref32_nomask), 4e-3)


def test_relative_position_buckets_match_transformers():
    from transformers import UMT5Config
    from transformers.models.umt5.modeling_umt5 import UMT5Attention
    from alg_b200.encoders import relative_position_buckets
    att = UMT5Attention(UMT5Config(d_model=64, d_kv=16, num_heads=4, relative_attention_num_buckets=32, relative_attention_max_distance=128),
                        has_relative_attention_bias=True)
    att.is_decoder = False
    n = 300
    rel = torch.arange(n)[None, :] - torch.arange(n)[:, None]
    want = att._relative_position_bucket(rel)
    got = relative_position_buckets(n, n, 32, 128)
    assert torch.equal(got[rel + n - 1], want)


@pytest.mark.parametrize("over,tol", [
    (dict(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=4, image_size=56, patch_size=14), 1e-4),
    (dict(hidden_size=96, intermediate_size=200, num_hidden_layers=2, num_attention_heads=2, image_size=28, patch_size=14, hidden_act="quick_gelu"), 1e-4),
    (dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=2, num_attention_heads=16, image_size=224, patch_size=14), 1e-4),
    # CLIP-ViT-L/14-336 geometry (LLaVA's tower): 577 tokens x head_dim 64 in fp32 does not fit one head in shared memory -> streamed keys
    (dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, image_size=336, patch_size=14, hidden_act="quick_gelu"), 1e-4),
])
def test_clip_vision_matches_transformers(over, tol):
    from transformers import CLIPVisionConfig, CLIPVisionModel
    from alg_b200 import encoders
    torch.manual_seed(3)
    cfg = dict(hidden_act="gelu", layer_norm_eps=1e-5, num_channels=3)
    cfg.update(over)
    hf = CLIPVisionModel(CLIPVisionConfig(**cfg)).eval()
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.copy_(1 + 0.1 * torch.randn_like(p))
            elif n.endswith("bias"):
                p.copy_(0.05 * torch.randn_like(p))
            elif p.dim() >= 2:
                p.copy_(torch.randn_like(p) * (p[0].numel() ** -0.5))
            else:
                p.copy_(0.5 * torch.randn_like(p))
    hf = hf.float().cuda()
    mine = encoders.CLIPVisionModel(**cfg).load_state_dict({k: v.detach().clone() for k, v in hf.state_dict().items()})
    px = torch.randn(2, 3, cfg["image_size"], cfg["image_size"], device="cuda")
    with torch.no_grad():
        ref = hf(pixel_values=px, output_hidden_states=True).hidden_states
    out = mine(pixel_values=px, output_hidden_states=True).hidden_states
    assert len(out) == len(ref) == cfg["num_hidden_layers"] + 1
    for i, (a, b) in enumerate(zip(out, ref)):
        assert a.dtype == torch.float32 and rel_l2(a, b) < tol, (i, rel_l2(a, b))


@pytest.mark.parametrize("over,eos_legacy", [
    (dict(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=4, vocab_size=1000, eos_token_id=2), True),
    (dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=2, num_attention_heads=12, vocab_size=49408, eos_token_id=49407), False),
])
def test_clip_text_matches_transformers(over, eos_legacy):
    """HunyuanVideo text_encoder_2 (hy:421-452): ids only (no attention mask), causal, pooled = the EOS token's final-LN state."""
    from transformers import CLIPTextConfig, CLIPTextModel
    from alg_b200 import encoders
    torch.manual_seed(7)
    cfg = dict(max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5, bos_token_id=0, pad_token_id=1)
    cfg.update(over)
    hf = CLIPTextModel(CLIPTextConfig(**cfg)).eval()
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.copy_(1 + 0.1 * torch.randn_like(p))
            elif n.endswith("bias"):
                p.copy_(0.05 * torch.randn_like(p))
            elif "embedding" in n:
                p.copy_(0.5 * torch.randn_like(p))
            else:
                p.copy_(torch.randn_like(p) * (p.shape[1] ** -0.5))
    hf = hf.float().cuda()
    mine = encoders.CLIPTextModel(**cfg).load_state_dict({k: v.detach().clone() for k, v in hf.state_dict().items()})
    g = torch.Generator().manual_seed(11)
    eos = cfg["eos_token_id"]
    hi = cfg["vocab_size"] - 1 if eos_legacy else eos  # legacy pooling = argmax of the ids: keep EOS the largest id only when it is
    ids = torch.randint(3, hi, (3, 77), generator=g)
    for b, n in enumerate((9, 77, 40)):  # prompt, EOS, then padding with the EOS id (CLIP tokenizers pad with <|endoftext|>)
        ids[b, n - 1:] = eos if not eos_legacy else 1
        if eos_legacy:
            ids[b, n - 1] = cfg["vocab_size"] - 1
    ids = ids.cuda()
    with torch.no_grad():
        ref = hf(ids, output_hidden_states=True)
    out = mine(ids, output_hidden_states=True)
    assert out.pooler_output.shape == ref.pooler_output.shape and out.pooler_output.dtype == torch.float32
    assert rel_l2(out.pooler_output, ref.pooler_output) < 1e-4, rel_l2(out.pooler_output, ref.pooler_output)
    assert rel_l2(out.last_hidden_state, ref.last_hidden_state) < 1e-4
    for i, (a, b) in enumerate(zip(out.hidden_states, ref.hidden_states)):
        assert rel_l2(a, b) < 1e-4, (i, rel_l2(a, b))
    assert mine(ids).hidden_states is None


def test_small_attention_masks_causal_and_fp32_against_torch():
    """alg_small_attention alone: causal flag (CLIP text towers), key-padding, head_dim 80, fp32 and bf16."""
    import torch.nn.functional as F
    from alg_b200.encoders import small_attention
    g = torch.Generator(device="cuda").manual_seed(0)
    for dt, tol in ((torch.float32, 2e-6), (torch.bfloat16, 2 ** -7)):
        B, L, H, D = 2, 77, 3, 80
        q, k, v = (torch.randn(B, L, H, D, generator=g, device="cuda").to(dt) for _ in range(3))
        valid = torch.tensor([77, 41], device="cuda", dtype=torch.int32)
        out = torch.empty_like(q)
        small_attention(q, k, v, out, batch=B, heads=H, head_dim=D, n_q=L, n_kv=L, q_bs=L * H * D, q_rs=H * D, k_bs=L * H * D,
                        k_rs=H * D, v_bs=L * H * D, v_rs=H * D, o_bs=L * H * D, o_rs=H * D, scale=D ** -0.5, kv_valid=valid, causal=True)
        mask = torch.ones(L, L, device="cuda", dtype=torch.bool).tril()[None, None] & (torch.arange(L, device="cuda")[None, :] < valid[:, None])[:, None, None, :]
        ref = F.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2), attn_mask=mask).transpose(1, 2)
        assert rel_l2(out, ref) < tol, (dt, rel_l2(out, ref))


def test_wan_pipeline_conditioning_through_native_encoders():
    """wan:185-234 end to end: tokenizer -> native UMT5 -> zero-padded prompt_embeds; image_processor -> native CLIP ->
    hidden_states[-2]; both against the same call sequence on the transformers modules."""
    from transformers import CLIPVisionConfig, CLIPVisionModel as HFCLIP, UMT5EncoderModel as HFUMT5
    from alg_b200 import encoders
    from alg_b200.pipeline_utils import SyntheticImageProcessor, SyntheticTokenizer
    from alg_b200.schedulers import UniPCMultistepScheduler
    from oracle.stub_vae import ArithVAE
    from pipeline_wan_image2video_lowpass import WanImageToVideoPipeline
    base, hf_t = _umt5("UMT5EncoderModel", dict(vocab_size=4096))
    hf_t = hf_t.to(torch.bfloat16).cuda()
    ccfg = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=4, image_size=224, patch_size=14)
    torch.manual_seed(5)
    hf_c = HFCLIP(CLIPVisionConfig(**ccfg)).eval().float().cuda()
    text = encoders.UMT5EncoderModel(**base).load_state_dict({k: v.clone() for k, v in hf_t.state_dict().items()})
    clip = encoders.CLIPVisionModel(**ccfg).load_state_dict({k: v.clone() for k, v in hf_c.state_dict().items()})
    tok, proc = SyntheticTokenizer(vocab_size=4096), SyntheticImageProcessor()
    from types import SimpleNamespace
    dummy = SimpleNamespace(config=SimpleNamespace(patch_size=(1, 2, 2)), dtype=torch.bfloat16, to=lambda *a, **k: None)
    pipes = [WanImageToVideoPipeline(tokenizer=tok, text_encoder=t, image_encoder=c, image_processor=proc, transformer=dummy,
                                     vae=ArithVAE("wan"), scheduler=UniPCMultistepScheduler()).to("cuda") for t, c in ((text, clip), (hf_t, hf_c))]
    prompts = ["a red bus turning a corner in the rain", "blurry"]
    image = torch.rand(1, 3, 96, 128)
    with torch.no_grad():
        mine = pipes[0]._get_t5_prompt_embeds(prompts, 2, 64), pipes[0].encode_image(image)
        ref = pipes[1]._get_t5_prompt_embeds(prompts, 2, 64), pipes[1].encode_image(image)
    assert mine[0].shape == ref[0].shape == (4, 64, 128) and rel_l2(mine[0], ref[0]) < 2e-2
    assert bool((mine[0][0, 10:] == 0).all())  # zero beyond the prompt length (wan:214-217)
    assert mine[1].shape == ref[1].shape == (1, 257, 128) and rel_l2(mine[1], ref[1]) < 1e-4


def test_from_pretrained_loads_native_encoders_from_a_snapshot(tmp_path):
    """run.py:46-61 on a LOCAL snapshot: transformer/ + text_encoder/ + image_encoder/ folders load into the native engines
    (diffusers / transformers parameter names, no key mapping); without the encoder folders from_pretrained refuses instead
    of silently pairing real DiT weights with stand-in conditioning."""
    from alg_b200 import checkpoint, encoders, wan
    from alg_b200.pipeline_utils import SyntheticImageProcessor, SyntheticTokenizer
    from oracle.stub_vae import ArithVAE
    from pipeline_wan_image2video_lowpass import WanImageToVideoPipeline
    import __graft_entry__ as G
    cfg, model, _, _ = G.tiny_problem("cuda")
    snap = str(tmp_path / "snap")
    checkpoint.save_transformer(snap, dict(wan.WAN_I2V_14B, **cfg), model.state_dict(), "WanTransformer3DModel")
    with pytest.raises(NotImplementedError, match="text_encoder"):
        WanImageToVideoPipeline.from_pretrained(snap, vae=ArithVAE("wan"))
    tcfg = dict(vocab_size=512, d_model=64, d_kv=32, d_ff=128, num_layers=2, num_heads=2, relative_attention_num_buckets=32,
                relative_attention_max_distance=128, layer_norm_epsilon=1e-6, model_type="umt5", feed_forward_proj="gated-gelu")
    text = encoders.UMT5EncoderModel.from_synthetic(seed=1, **{k: v for k, v in tcfg.items() if k != "model_type"})
    ccfg = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2, image_size=224, patch_size=14,
                hidden_act="gelu", layer_norm_eps=1e-5, num_channels=3)
    clip = encoders.CLIPVisionModel.from_synthetic(seed=2, **ccfg)
    checkpoint.save_component(snap, "text_encoder", tcfg, text.state_dict())
    checkpoint.save_component(snap, "image_encoder", ccfg, clip.state_dict())
    pipe = WanImageToVideoPipeline.from_pretrained(snap, vae=ArithVAE("wan"), tokenizer=SyntheticTokenizer(vocab_size=512),
                                                   image_processor=SyntheticImageProcessor()).to("cuda")
    assert isinstance(pipe.text_encoder, encoders.UMT5EncoderModel) and isinstance(pipe.image_encoder, encoders.CLIPVisionModel)
    tok = pipe.tokenizer(["a red bus"], max_length=32)
    a = pipe.text_encoder(tok.input_ids.cuda(), tok.attention_mask.cuda()).last_hidden_state
    b = text(tok.input_ids.cuda(), tok.attention_mask.cuda()).last_hidden_state
    assert torch.equal(a, b)
    img = torch.rand(1, 3, 64, 64)
    assert torch.equal(pipe.encode_image(img), clip(pixel_values=SyntheticImageProcessor()(images=img).pixel_values.cuda()).hidden_states[-2])
