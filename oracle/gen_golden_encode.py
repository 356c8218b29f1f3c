"""Golden vectors of HunyuanVideo's prompt-encoding host logic, produced by the UNMODIFIED reference pipeline.

TEST INFRASTRUCTURE.  Run in the build container (it needs /root/reference):

    python oracle/gen_golden_encode.py            # writes tests/golden/encode_hunyuan.npz

``oracle/refshim`` makes the reference module importable (see gen_golden_loops.py); ``oracle/stub_text.py`` supplies the
third-party objects (tokenizers, LLaVA / CLIP encoders, image processor) as closed-form stand-ins.  What runs is the
reference's own ``encode_prompt`` -> ``_get_llama_prompt_embeds`` (template formatting, ``_expand_input_ids_with_image_tokens``,
hidden-state selection, template / assistant-header cropping, image-slot interleaving) and ``_get_clip_prompt_embeds``
(hy:107-149, 282-492)."""
from __future__ import annotations

import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path[:0] = [os.path.join(HERE, "refshim"), REF]
sys.path.append(ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import pipeline_hunyuan_video_image2video_lowpass as ref_hy  # noqa: E402
from oracle.stub_text import ClosedFormClip, ClosedFormLlava, PixelProcessor, TemplateTokenizer, WordTokenizer  # noqa: E402
from oracle.stub_vae import ArithVAE  # noqa: E402

assert ref_hy.__file__.startswith(REF)

CASES = {
    # name: (prompts, template overrides, max_sequence_length, image_embed_interleave)
    "default": (["a red bus turning a corner in the rain"], {}, 40, 2),
    "batch_two_interleave4": (["a cat", "two dogs running on a beach at sunset with waves"], {}, 32, 4),
    "truncated_prompt": ([" ".join(f"w{i}" for i in range(80))], {}, 24, 2),   # only 3 double-return tokens survive (hy:346-351)
    "no_interleave": (["a kite"], {}, 16, 0),
    "crop_start_from_tokenizer": (["a small boat"], {"crop_start": None}, 20, 2),
}


def template_for(tok: TemplateTokenizer, image_emb_len: int, overrides: dict) -> dict:
    t = dict(ref_hy.DEFAULT_PROMPT_TEMPLATE)
    system = t["template"].split("<|start_header_id|>user")[0]
    t["crop_start"] = len(tok.encode(system))           # tokens of the system block (the real tokenizer gives 103)
    t["image_emb_len"], t["image_emb_start"], t["image_emb_end"] = image_emb_len, 5, 5 + image_emb_len
    t.update(overrides)
    return t


def main():
    out, meta = {}, {}
    for name, (prompts, over, max_len, interleave) in CASES.items():
        # crop_start=None: the reference tokenizes the bare template padded to the tokenizer's model_max_length (hy:299-310)
        tok = TemplateTokenizer(model_max_length=96 if over.get("crop_start", 0) is None else 160)
        tmpl = template_for(tok, 12, over)
        pipe = ref_hy.HunyuanVideoImageToVideoPipeline(text_encoder=ClosedFormLlava(), tokenizer=tok, transformer=None,
                                                       vae=ArithVAE("hunyuan"), scheduler=None, text_encoder_2=ClosedFormClip(),
                                                       tokenizer_2=WordTokenizer(), image_processor=PixelProcessor())
        image = torch.linspace(0, 1, 3 * 16 * 16).view(3, 16, 16)
        embeds, pooled, mask = pipe.encode_prompt(image=image, prompt=prompts, prompt_template=tmpl, device=torch.device("cpu"),
                                                  max_sequence_length=max_len, image_embed_interleave=interleave)
        out[f"{name}.embeds"], out[f"{name}.pooled"], out[f"{name}.mask"] = embeds.numpy(), pooled.numpy(), mask.numpy()
        meta[name] = dict(prompts=prompts, template=tmpl, max_sequence_length=max_len, image_embed_interleave=interleave,
                          model_max_length=tok.model_max_length)
        print(name, tuple(embeds.shape), tuple(pooled.shape), tuple(mask.shape), int(mask.sum()))
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "encode_hunyuan.npz"), **out)


if __name__ == "__main__":
    main()
