"""Native LLaVA-Llama-3 prompt encoder (alg_b200/llava.py over the C ABI) against the REAL transformers
``LlavaForConditionalGeneration`` with seeded random weights (transformers is installed, so this row's oracle is upstream code).

Tolerance: the engine computes in float32 (bf16x3 tensor-core linears), the reference runs this encoder in float16
(run.py:76-80).  Required: every hidden state within 2e-4 relative L2 of transformers' float32 forward on the same
fp16-rounded weights, and closer to it than transformers' own float16 forward is."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TEXT = dict(vocab_size=600, hidden_size=128, intermediate_size=352, num_hidden_layers=4, num_attention_heads=4,
            num_key_value_heads=2, rms_norm_eps=1e-5, max_position_embeddings=2048)
VISION = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=2, image_size=56, patch_size=14,
              num_channels=3, hidden_act="quick_gelu", layer_norm_eps=1e-5)
IMG, PAD = 590, 591


def _hf(text=TEXT, vision=VISION, seed=0):
    from transformers import CLIPVisionConfig, LlamaConfig, LlavaConfig, LlavaForConditionalGeneration
    torch.manual_seed(seed)
    cfg = LlavaConfig(vision_config=CLIPVisionConfig(**vision), text_config=LlamaConfig(rope_theta=500000.0, **text),
                      image_token_index=IMG, pad_token_id=PAD, vision_feature_layer=-2, vision_feature_select_strategy="default",
                      projector_hidden_act="gelu")
    hf = LlavaForConditionalGeneration(cfg).eval()
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.copy_(1 + 0.1 * torch.randn_like(p))
            elif n.endswith("bias"):
                p.copy_(0.05 * torch.randn_like(p))
            elif "embed" in n:
                p.copy_(0.5 * torch.randn_like(p))
            elif p.dim() >= 2:
                p.copy_(torch.randn_like(p) * (p[0].numel() ** -0.5))
            p.copy_(p.half().float())  # the checkpoint is fp16
    return hf.cuda()


def _inputs(B, L, lens, n_img, vocab=580, seed=1):
    """Expanded inputs like hy:107-149 produces: BOS-ish prefix, a block of image slots at [5, 5 + n_img), prompt, right padding."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, vocab, (B, L), generator=g)
    ids[:, 5:5 + n_img] = IMG
    mask = torch.zeros(B, L, dtype=torch.int64)
    for b, n in enumerate(lens):
        mask[b, :n] = 1
        ids[b, n:] = PAD
    pos = (mask.cumsum(-1) - 1).masked_fill_(mask == 0, 1)
    px = torch.randn(B, 3, VISION["image_size"], VISION["image_size"], generator=g)
    return ids.cuda(), mask.cuda(), pos.cuda(), px.cuda()


def test_llava_hidden_states_match_transformers():
    from alg_b200 import llava
    hf = _hf()
    mine = llava.LlavaForConditionalGeneration(text_config=dict(TEXT, rope_theta=500000.0), vision_config=VISION, image_token_index=IMG,
                                               pad_token_id=PAD).load_state_dict({k: v.detach().clone() for k, v in hf.state_dict().items()})
    n_img = (VISION["image_size"] // VISION["patch_size"]) ** 2
    ids, mask, pos, px = _inputs(2, 150, (150, 97), n_img)
    with torch.no_grad():
        ref32 = hf(input_ids=ids, attention_mask=mask, position_ids=pos, pixel_values=px, output_hidden_states=True).hidden_states
        hf16 = hf.half()
        ref16 = hf16(input_ids=ids, attention_mask=mask, position_ids=pos, pixel_values=px.half(), output_hidden_states=True).hidden_states
    out = mine(input_ids=ids, attention_mask=mask, position_ids=pos, pixel_values=px, output_hidden_states=True).hidden_states
    assert len(out) == len(ref32) == TEXT["num_hidden_layers"] + 1
    assert mine.config.image_token_index == IMG and mine.config.pad_token_id == PAD
    for i, (a, r32, r16) in enumerate(zip(out, ref32, ref16)):
        assert a.dtype == torch.float32 and a.shape == r32.shape
        for b, n in enumerate((150, 97)):  # attended rows; padded rows are cropped or masked downstream (hy:362-391)
            e = rel_l2(a[b, :n], r32[b, :n])
            assert e < 2e-4, (i, b, e)
            assert e <= rel_l2(r16[b, :n].float(), r32[b, :n]) + 1e-6, (i, b)
    # padded query rows are well defined too (causal + key padding, position 1): compare them where transformers is finite
    pad_rows = out[-3][1, 97:], ref32[-3][1, 97:]
    if bool(torch.isfinite(pad_rows[1]).all()):
        assert rel_l2(*pad_rows) < 2e-4


def test_llava_text_only_general_mask_and_old_checkpoint_names():
    """No pixel_values; a padding mask that is NOT a suffix (general key mask path); 4.48-style parameter names load."""
    from alg_b200 import llava
    hf = _hf(seed=3)
    old = {}
    for k, v in hf.state_dict().items():  # transformers 4.48 layout, as saved in the HunyuanVideo snapshots
        if k.startswith("model.language_model."):
            k = "language_model.model." + k[len("model.language_model."):]
        elif k.startswith("model."):
            k = k[len("model."):]
        elif k.startswith("lm_head."):
            k = "language_model." + k
        old[k] = v.detach().clone()
    mine = llava.LlavaForConditionalGeneration(text_config=dict(TEXT, rope_theta=500000.0), vision_config=VISION, image_token_index=IMG,
                                               pad_token_id=PAD).load_state_dict(old)
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(3, 580, (2, 70), generator=g).cuda()
    mask = torch.ones(2, 70, dtype=torch.int64)
    mask[0, :6] = 0       # left padding
    mask[1, 30:41] = 0    # a hole
    mask = mask.cuda()
    pos = (mask.cumsum(-1) - 1).masked_fill_(mask == 0, 1)
    with torch.no_grad():
        ref = hf(input_ids=ids, attention_mask=mask, position_ids=pos, output_hidden_states=True).hidden_states
    out = mine(input_ids=ids, attention_mask=mask, position_ids=pos).hidden_states
    keep = mask.bool()
    for i, (a, r) in enumerate(zip(out, ref)):
        assert rel_l2(a[keep], r[keep]) < 2e-4, (i, rel_l2(a[keep], r[keep]))
    # state_dict round trip (fp16 weights split exactly into bf16 hi + lo)
    sd = mine.state_dict()
    for k, v in hf.state_dict().items():
        if k in sd:
            assert torch.equal(sd[k], v.float()), k


def test_streamed_attention_gqa_long_prompt_against_torch():
    """alg_small_attention, streamed-key variant alone: 934 keys (the LLaVA prompt length), 8 query heads on 2 K/V heads,
    causal + right padding, fp32 and bf16; and a general key mask."""
    import torch.nn.functional as F
    from alg_b200.encoders import small_attention
    g = torch.Generator(device="cuda").manual_seed(0)
    B, L, H, Hk, D = 2, 934, 8, 2, 128
    for dt, tol in ((torch.float32, 3e-6), (torch.bfloat16, 2 ** -7)):
        q = torch.randn(B, L, H, D, generator=g, device="cuda").to(dt)
        k, v = (torch.randn(B, L, Hk, D, generator=g, device="cuda").to(dt) for _ in range(2))
        valid = torch.tensor([934, 611], device="cuda", dtype=torch.int32)
        out = torch.empty_like(q)
        small_attention(q, k, v, out, batch=B, heads=H, head_dim=D, n_q=L, n_kv=L, q_bs=L * H * D, q_rs=H * D, k_bs=L * Hk * D,
                        k_rs=Hk * D, v_bs=L * Hk * D, v_rs=Hk * D, o_bs=L * H * D, o_rs=H * D, scale=D ** -0.5, kv_valid=valid,
                        causal=True, kv_group=H // Hk)
        m = torch.ones(L, L, device="cuda", dtype=torch.bool).tril()[None, None] & (torch.arange(L, device="cuda")[None, :] < valid[:, None])[:, None, None, :]
        kk, vv = (t.float().transpose(1, 2).repeat_interleave(H // Hk, dim=1) for t in (k, v))
        ref = F.scaled_dot_product_attention(q.float().transpose(1, 2), kk, vv, attn_mask=m).transpose(1, 2)
        assert rel_l2(out, ref) < tol, (dt, rel_l2(out, ref))
    km = (torch.rand(B, L, generator=g, device="cuda") > 0.3)
    km[:, 0] = True
    q, k, v = (torch.randn(B, L, H, D, generator=g, device="cuda") for _ in range(3))
    out = torch.empty_like(q)
    small_attention(q, k, v, out, batch=B, heads=H, head_dim=D, n_q=L, n_kv=L, q_bs=L * H * D, q_rs=H * D, k_bs=L * H * D, k_rs=H * D,
                    v_bs=L * H * D, v_rs=H * D, o_bs=L * H * D, o_rs=H * D, scale=D ** -0.5, causal=False, key_mask=km.to(torch.uint8))
    ref = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), attn_mask=km[:, None, None, :]).transpose(1, 2)
    assert rel_l2(out, ref) < 3e-6


def test_hunyuan_pipeline_conditioning_through_native_encoders():
    """hy:282-452 end to end on the GPU: template tokenizer -> native LLaVA / HF LLaVA, CLIP tokenizer -> native / HF CLIP text."""
    from transformers import CLIPTextConfig, CLIPTextModel as HFCLIPText
    from alg_b200 import encoders, llava
    from oracle.stub_text import IMAGE_TOKEN, PAD as TPAD, PixelProcessor, TemplateTokenizer, WordTokenizer
    from oracle.stub_vae import ArithVAE
    import pipeline_hunyuan_video_image2video_lowpass as P
    text = dict(TEXT, vocab_size=128320, hidden_size=64, intermediate_size=128, num_hidden_layers=4)
    from transformers import CLIPVisionConfig, LlamaConfig, LlavaConfig, LlavaForConditionalGeneration
    torch.manual_seed(9)
    hf = LlavaForConditionalGeneration(LlavaConfig(vision_config=CLIPVisionConfig(**VISION), text_config=LlamaConfig(rope_theta=500000.0, **text),
                                                   image_token_index=IMAGE_TOKEN, pad_token_id=TPAD)).eval().float().cuda()
    ccfg = dict(vocab_size=49408, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5, eos_token_id=49407, bos_token_id=49406, pad_token_id=49407)
    hf_c = HFCLIPText(CLIPTextConfig(**ccfg)).eval().float().cuda()
    mine = llava.LlavaForConditionalGeneration(text_config=dict(text, rope_theta=500000.0), vision_config=VISION,
                                               image_token_index=IMAGE_TOKEN, pad_token_id=TPAD).load_state_dict(
        {k: v.detach().clone() for k, v in hf.state_dict().items()})
    mine_c = encoders.CLIPTextModel(**ccfg).load_state_dict({k: v.detach().clone() for k, v in hf_c.state_dict().items()})

    class Proc(PixelProcessor):  # the vision tower wants image_size x image_size
        def __call__(self, image, return_tensors="pt", **kw):
            x = torch.as_tensor(image).float()
            x = x[None] if x.dim() == 3 else x
            from types import SimpleNamespace
            return SimpleNamespace(pixel_values=torch.nn.functional.interpolate(x, size=(VISION["image_size"],) * 2, mode="bilinear"))

    tok = TemplateTokenizer()
    n_img = (VISION["image_size"] // VISION["patch_size"]) ** 2
    tmpl = dict(P.DEFAULT_PROMPT_TEMPLATE)
    tmpl["crop_start"] = len(tok.encode(tmpl["template"].split("<|start_header_id|>user")[0]))
    tmpl["image_emb_len"], tmpl["image_emb_end"] = n_img, 5 + n_img
    res = []
    for te, te2 in ((mine, mine_c), (hf, hf_c)):
        pipe = P.HunyuanVideoImageToVideoPipeline(text_encoder=te, tokenizer=tok, transformer=None, vae=ArithVAE("hunyuan"),
                                                  scheduler=None, text_encoder_2=te2, tokenizer_2=WordTokenizer(), image_processor=Proc())
        with torch.no_grad():
            res.append(pipe.encode_prompt(image=torch.rand(3, 64, 96, generator=torch.Generator().manual_seed(2)),
                                          prompt=["a red bus turning a corner"], prompt_template=tmpl,
                                          device=torch.device("cuda"), max_sequence_length=32))
            # one image for two prompts: transformers refuses (image slots != feature rows), and so does the native encoder
            with pytest.raises(ValueError, match="do not match"):
                pipe.encode_prompt(image=torch.rand(3, 64, 96), prompt=["a red bus", "rain"], prompt_template=tmpl,
                                   device=torch.device("cuda"), max_sequence_length=32)
    (e0, p0, m0), (e1, p1, m1) = res
    assert e0.shape == e1.shape and torch.equal(m0, m1)
    keep = m0.bool()
    assert rel_l2(e0[keep], e1[keep]) < 2e-4 and rel_l2(p0, p1) < 1e-4


def test_hunyuan_pipeline_end_to_end_all_native_vs_upstream_modules():
    """run.py's HunyuanVideo path on tiny shapes -- image + prompt -> template tokenizer -> LLaVA, CLIP text, VAE encode of the
    first frame, ALG loop (true CFG: three-pass then two-pass steps), VAE decode, frames: once with every network native, once
    with the transformers modules and the oracle VAE around the same native DiT."""
    from types import SimpleNamespace
    import numpy as np
    from transformers import CLIPTextConfig, CLIPTextModel as HFCLIPText, CLIPVisionConfig, LlamaConfig, LlavaConfig
    from transformers import LlavaForConditionalGeneration as HFLlava
    from alg_b200 import encoders, hunyuan, llava
    from alg_b200.schedulers import FlowMatchEulerDiscreteScheduler
    from alg_b200.vae_hunyuan import AutoencoderKLHunyuanVideo
    from oracle import hunyuan_vae_oracle as V
    from oracle.stub_text import IMAGE_TOKEN, PAD as TPAD, TemplateTokenizer, WordTokenizer
    import pipeline_hunyuan_video_image2video_lowpass as P
    dit = hunyuan.HunyuanVideoTransformer3DModel.from_synthetic(seed=0, device="cuda", num_attention_heads=2, attention_head_dim=128,
                                                                num_layers=2, num_single_layers=2, num_refiner_layers=1,
                                                                text_embed_dim=64, pooled_projection_dim=32)
    text = dict(TEXT, vocab_size=128320, hidden_size=64, intermediate_size=128, num_hidden_layers=4)
    torch.manual_seed(9)
    hf = HFLlava(LlavaConfig(vision_config=CLIPVisionConfig(**VISION), text_config=LlamaConfig(rope_theta=500000.0, **text),
                             image_token_index=IMAGE_TOKEN, pad_token_id=TPAD)).eval().float().cuda()
    ccfg = dict(vocab_size=49408, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2,
                max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5, eos_token_id=49407, bos_token_id=49406, pad_token_id=49407)
    hf_c = HFCLIPText(CLIPTextConfig(**ccfg)).eval().float().cuda()
    mine = llava.LlavaForConditionalGeneration(text_config=dict(text, rope_theta=500000.0), vision_config=VISION, image_token_index=IMAGE_TOKEN,
                                               pad_token_id=TPAD).load_state_dict({k: v.detach().clone() for k, v in hf.state_dict().items()})
    mine_c = encoders.CLIPTextModel(**ccfg).load_state_dict({k: v.detach().clone() for k, v in hf_c.state_dict().items()})
    vcfg = dict(V.HUNYUAN_VAE, block_out_channels=(32, 64, 64, 64))
    sd = V.make_weights(vcfg, seed=4, device="cuda")
    vae = AutoencoderKLHunyuanVideo(block_out_channels=(32, 64, 64, 64)).load_state_dict({k: v.clone() for k, v in sd.items()})
    sd64 = {k: v.double() for k, v in sd.items()}

    class OracleVAE:
        dtype = torch.float32
        config = SimpleNamespace(**vcfg)
        temporal_compression_ratio, spatial_compression_ratio = 4, 8

        def encode(self, x):
            m = V.encode_moments(x.double(), sd64, vcfg, torch.float64).float()
            return SimpleNamespace(latent_dist=SimpleNamespace(mode=lambda: m[:, :16], sample=lambda generator=None: m[:, :16]))

        def decode(self, z, return_dict=True):
            v = V.decode(z.double(), sd64, vcfg, torch.float64).float()
            return SimpleNamespace(sample=v) if return_dict else (v,)

    class Proc:
        def __call__(self, image, return_tensors="pt", **kw):
            x = torch.as_tensor(np.asarray(image), dtype=torch.float32)
            x = x.permute(2, 0, 1)[None] / 255.0 if x.dim() == 3 and x.shape[-1] == 3 else (x[None] if x.dim() == 3 else x)
            return SimpleNamespace(pixel_values=torch.nn.functional.interpolate(x, size=(VISION["image_size"],) * 2, mode="bilinear"))

    tok = TemplateTokenizer()
    n_img = (VISION["image_size"] // VISION["patch_size"]) ** 2
    tmpl = dict(P.DEFAULT_PROMPT_TEMPLATE)
    tmpl["crop_start"] = len(tok.encode(tmpl["template"].split("<|start_header_id|>user")[0]))
    tmpl["image_emb_len"], tmpl["image_emb_end"] = n_img, 5 + n_img
    alg = dict(use_low_pass_guidance=True, lp_filter_type="down_up", lp_filter_in_latent=True, lp_blur_sigma=15.0,
               lp_blur_kernel_size=0.02734375, lp_resize_factor=0.625, lp_strength_schedule_type="interval",
               schedule_blur_kernel_size=False, schedule_interval_start_time=0.0, schedule_interval_end_time=0.4,
               schedule_linear_start_weight=1.0, schedule_linear_end_weight=0.0, schedule_linear_end_time=0.5,
               schedule_exp_decay_rate=10.0)
    image = torch.rand(1, 3, 64, 96, generator=torch.Generator().manual_seed(3))
    frames = []
    for te, te2, v in ((mine, mine_c, vae), (hf, hf_c, OracleVAE())):
        pipe = P.HunyuanVideoImageToVideoPipeline(text_encoder=te, tokenizer=tok, transformer=dit, vae=v,
                                                  scheduler=FlowMatchEulerDiscreteScheduler(shift=7.0), text_encoder_2=te2,
                                                  tokenizer_2=WordTokenizer(), image_processor=Proc()).to("cuda")
        out = pipe(image=image, prompt="a red bus turning a corner in the rain", negative_prompt="blurry", height=64, width=96,
                   num_frames=5, num_inference_steps=3, guidance_scale=6.0, true_cfg_scale=4.0, prompt_template=tmpl,
                   max_sequence_length=32, generator=torch.Generator(device="cuda").manual_seed(42), output_type="np", **alg)
        frames.append(np.asarray(out.frames))
    a, b = frames
    assert a.shape == b.shape == (1, 5, 64, 96, 3) and np.isfinite(a).all() and a.min() >= 0.0 and a.max() <= 1.0 and a.std() > 1e-3
    err = np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64))
    assert err < 2e-2, err


def test_llama_stack_at_true_width_two_layers():
    """Llama-3-8B geometry (hidden 4096, 32 query heads on 8 K/V heads of 128, SwiGLU 14336, rope_theta 5e5) on two layers and a
    900-token prompt -- the head_dim / group / key-count combination the HunyuanVideo encoder really runs."""
    from transformers import CLIPVisionConfig, LlamaConfig, LlavaConfig, LlavaForConditionalGeneration
    from alg_b200 import llava
    text = dict(vocab_size=2048, hidden_size=4096, intermediate_size=14336, num_hidden_layers=2, num_attention_heads=32,
                num_key_value_heads=8, rms_norm_eps=1e-5, max_position_embeddings=8192)
    torch.manual_seed(11)
    hf = LlavaForConditionalGeneration(LlavaConfig(vision_config=CLIPVisionConfig(**VISION), text_config=LlamaConfig(rope_theta=500000.0, **text),
                                                   image_token_index=2040, pad_token_id=2041)).eval()
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.copy_(1 + 0.1 * torch.randn_like(p))
            elif "embed" in n:
                p.copy_(0.5 * torch.randn_like(p))
            elif p.dim() >= 2:
                p.copy_(torch.randn_like(p) * (p[0].numel() ** -0.5))
            p.copy_(p.half().float())
    hf = hf.cuda()
    mine = llava.LlavaForConditionalGeneration(text_config=dict(text, rope_theta=500000.0), vision_config=VISION, image_token_index=2040,
                                               pad_token_id=2041).load_state_dict({k: v.detach().clone() for k, v in hf.state_dict().items()})
    L, n = 900, 731
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(3, 2000, (1, L), generator=g)
    mask = torch.zeros(1, L, dtype=torch.int64)
    mask[:, :n] = 1
    ids[:, n:] = 2041
    pos = (mask.cumsum(-1) - 1).masked_fill_(mask == 0, 1)
    ids, mask, pos = ids.cuda(), mask.cuda(), pos.cuda()
    with torch.no_grad():
        ref = hf(input_ids=ids, attention_mask=mask, position_ids=pos, output_hidden_states=True).hidden_states
        ref16 = hf.half()(input_ids=ids, attention_mask=mask, position_ids=pos, output_hidden_states=True).hidden_states
    out = mine(input_ids=ids, attention_mask=mask, position_ids=pos).hidden_states
    for i, (a, r, r16) in enumerate(zip(out, ref, ref16)):
        e = rel_l2(a[0, :n], r[0, :n])
        assert e < 2e-4 and e <= rel_l2(r16[0, :n].float(), r[0, :n]) + 1e-6, (i, e)
