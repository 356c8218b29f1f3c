"""TEST INFRASTRUCTURE: deterministic stand-ins with the transformers call surface of HunyuanVideo's prompt stack
(tokenizer + LLaVA encoder + image processor, tokenizer_2 + CLIP text model), used to pin the FIRST-PARTY host logic of
``_get_llama_prompt_embeds`` / ``_get_clip_prompt_embeds`` / ``encode_prompt`` (hy:107-149, 282-492): the same stubs drive the
unmodified reference (``oracle/gen_golden_encode.py``) and the repo's pipeline (``tests/test_encode_prompt_golden.py``).

The "encoder" returns hidden states that are a closed-form function of what it was handed (ids, position ids, mask, slot
index, pixel mean), so any difference in expansion / cropping / interleaving shows up exactly."""
from __future__ import annotations

import re
import zlib
from types import SimpleNamespace

import torch

BOS, START_HEADER, END_HEADER, EOT, DOUBLE_RETURN = 128000, 128006, 128007, 128009, 271
IMAGE_TOKEN, PAD = 128257, 128258
_SPECIAL = {"<|start_header_id|>": START_HEADER, "<|end_header_id|>": END_HEADER, "<|eot_id|>": EOT, "<image>": IMAGE_TOKEN,
            "\n\n": DOUBLE_RETURN}


class _Enc(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class TemplateTokenizer:
    """Llama-3-like: BOS, the five special strings above as single tokens, every other whitespace-separated word one hashed id;
    right padding with PAD.  ``padding="max_length"`` without ``max_length`` pads to ``model_max_length``."""

    def __init__(self, model_max_length: int = 160):
        self.model_max_length = model_max_length
        self.pad_token_id = PAD

    def encode(self, text: str):
        ids = [BOS]
        for piece in re.split(r"(<\|start_header_id\|>|<\|end_header_id\|>|<\|eot_id\|>|<image>|\n\n)", text):
            if piece in _SPECIAL:
                ids.append(_SPECIAL[piece])
            else:
                ids += [1000 + zlib.crc32(w.encode()) % 100000 for w in piece.split()]
        return ids

    def __call__(self, prompt, max_length=None, padding="max_length", truncation=False, return_tensors="pt",
                 return_attention_mask=True, **kw):
        prompt = [prompt] if isinstance(prompt, str) else list(prompt)
        rows = [self.encode(p) for p in prompt]
        if truncation and max_length is not None:
            rows = [r[:max_length] for r in rows]
        width = (max_length or self.model_max_length) if padding == "max_length" else max(len(r) for r in rows)
        ids = torch.full((len(rows), width), PAD, dtype=torch.int64)
        mask = torch.zeros(len(rows), width, dtype=torch.int64)
        for b, r in enumerate(rows):
            ids[b, :len(r)] = torch.tensor(r)
            mask[b, :len(r)] = 1
        out = _Enc(input_ids=ids)
        if return_attention_mask:
            out["attention_mask"] = mask
        return out


class PixelProcessor:
    """``CLIPImageProcessor``-like: returns the image as ``pixel_values`` [1, 3, 8, 8] (mean-pooled)."""

    def __call__(self, image, return_tensors="pt", **kw):
        x = torch.as_tensor(image).float()
        x = x[None] if x.dim() == 3 else x
        return _Enc(pixel_values=torch.nn.functional.adaptive_avg_pool2d(x, 8))


class ClosedFormLlava:
    """hidden_states[k][b, i, :] = [input_ids, position_ids, attention_mask, i, mean(pixel_values), k] (+ zero padding to dim)."""

    def __init__(self, dim: int = 8, n_states: int = 5, dtype=torch.float32):
        self.config = SimpleNamespace(image_token_index=IMAGE_TOKEN, pad_token_id=PAD)
        self.dtype, self.dim, self.n_states = dtype, dim, n_states

    def __call__(self, input_ids=None, attention_mask=None, position_ids=None, pixel_values=None, output_hidden_states=True, **kw):
        B, L = input_ids.shape
        base = torch.zeros(B, L, self.dim, dtype=torch.float64)
        base[..., 0] = input_ids.double()
        base[..., 1] = position_ids.double()
        base[..., 2] = attention_mask.double()
        base[..., 3] = torch.arange(L).double()[None]
        base[..., 4] = pixel_values.double().mean()
        states = []
        for k in range(self.n_states):
            s = base.clone()
            s[..., 5] = k
            states.append(s.to(self.dtype))
        return SimpleNamespace(hidden_states=tuple(states))


class WordTokenizer:
    """CLIP-tokenizer-like: <|startoftext|> 49406, hashed words, <|endoftext|> 49407 (also the pad token)."""

    def __call__(self, prompt, padding="max_length", max_length=None, truncation=False, return_tensors="pt", **kw):
        prompt = [prompt] if isinstance(prompt, str) else list(prompt)
        rows = [[49406] + [1000 + zlib.crc32(w.encode()) % 40000 for w in p.split()] + [49407] for p in prompt]
        if truncation and max_length is not None:
            rows = [r[:max_length - 1] + [49407] if len(r) > max_length else r for r in rows]
        width = max_length if padding == "max_length" else max(len(r) for r in rows)
        ids = torch.full((len(rows), width), 49407, dtype=torch.int64)
        for b, r in enumerate(rows):
            ids[b, :len(r)] = torch.tensor(r)
        return _Enc(input_ids=ids)

    def batch_decode(self, ids):
        return [" ".join(str(int(t)) for t in row) for row in ids]


class ClosedFormClip:
    """pooler_output[b] = [first EOS position, sum of ids, number of tokens before EOS, 0...]."""

    def __init__(self, dim: int = 6, dtype=torch.float32):
        self.dtype, self.dim = dtype, dim

    def __call__(self, input_ids, output_hidden_states=False, **kw):
        B, L = input_ids.shape
        out = torch.zeros(B, self.dim, dtype=torch.float64)
        eos = (input_ids == 49407).int().argmax(dim=-1)
        out[:, 0] = eos.double()
        out[:, 1] = input_ids.double().sum(dim=-1)
        out[:, 2] = L
        return SimpleNamespace(pooler_output=out.to(self.dtype))
