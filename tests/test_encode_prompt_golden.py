"""HunyuanVideo prompt-encoding host logic (hy:107-149, 282-492) against vectors produced by the UNMODIFIED reference
(``oracle/gen_golden_encode.py``): same closed-form tokenizer / encoder stand-ins (``oracle/stub_text.py``) on both sides, so
template formatting, ``<image>`` expansion, position ids, hidden-state selection, template / assistant-header cropping and
image-slot interleaving must agree exactly.  CPU only: the method under test just sequences calls on the objects it is given."""
import json
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "encode_hunyuan.npz")


def _cases():
    z = np.load(GOLD)
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


@pytest.mark.parametrize("name", ["default", "batch_two_interleave4", "truncated_prompt", "no_interleave", "crop_start_from_tokenizer"])
def test_encode_prompt_matches_reference(name):
    from oracle.stub_text import ClosedFormClip, ClosedFormLlava, PixelProcessor, TemplateTokenizer, WordTokenizer
    from oracle.stub_vae import ArithVAE
    from pipeline_hunyuan_video_image2video_lowpass import HunyuanVideoImageToVideoPipeline
    z, meta = _cases()
    m = meta[name]
    pipe = HunyuanVideoImageToVideoPipeline(text_encoder=ClosedFormLlava(), tokenizer=TemplateTokenizer(m["model_max_length"]),
                                            transformer=None, vae=ArithVAE("hunyuan"), scheduler=None,
                                            text_encoder_2=ClosedFormClip(), tokenizer_2=WordTokenizer(),
                                            image_processor=PixelProcessor())
    image = torch.linspace(0, 1, 3 * 16 * 16).view(3, 16, 16)
    embeds, pooled, mask = pipe.encode_prompt(image=image, prompt=m["prompts"], prompt_template=m["template"],
                                              device=torch.device("cpu"), max_sequence_length=m["max_sequence_length"],
                                              image_embed_interleave=m["image_embed_interleave"])
    assert np.array_equal(embeds.numpy(), z[f"{name}.embeds"])
    assert np.array_equal(pooled.numpy(), z[f"{name}.pooled"])
    assert np.array_equal(mask.numpy(), z[f"{name}.mask"]) and mask.dtype == torch.int64


def test_expand_input_ids_positions_and_mask():
    """hy:107-149 on a hand-checkable row: one <image> at index 2 expands to 4 slots; pads stay masked with position 1."""
    from pipeline_hunyuan_video_image2video_lowpass import _expand_input_ids_with_image_tokens
    IMG, PAD = 9, 0
    ids = torch.tensor([[5, 6, IMG, 7, 8, PAD, PAD]])
    out = _expand_input_ids_with_image_tokens(ids, (ids != PAD).long(), 7, IMG, 4, 2, 6, PAD)
    assert out["input_ids"].tolist() == [[5, 6, IMG, IMG, IMG, IMG, 7, 8, PAD, PAD]]
    assert out["attention_mask"].tolist() == [[1, 1, 1, 1, 1, 1, 1, 1, 0, 0]]
    assert out["position_ids"].tolist() == [[0, 1, 2, 3, 4, 5, 6, 7, 1, 1]]
