"""Minimal stand-ins for the ``diffusers`` plumbing the three reference pipelines inherit.

``diffusers`` (requirements.txt:13) is not installable offline, so the re-hosted pipelines cannot subclass
``DiffusionPipeline``; this module carries the few pieces of it the hot path touches -- module registry + ``.to()``,
``_execution_device``, ``progress_bar``, ``randn_tensor``, ``VideoProcessor``, output dataclasses -- plus
SYNTHETIC encoders / VAEs (true latent geometry, seeded weights) so ``run.py`` works end to end without
checkpoints.  Nothing here is on the per-step path; the VAE and the conditioning encoders are SURVEY section 8(f)
"next" items and are clearly not the real networks.
"""
from __future__ import annotations

import contextlib
import os
import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import List, Optional, Union

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------
def randn_tensor(shape, generator=None, device=None, dtype=None):
    """diffusers.utils.torch_utils.randn_tensor: CPU generators draw on the CPU and move to ``device``."""
    device = torch.device(device or "cpu")
    if isinstance(generator, list):
        shape1 = (1,) + tuple(shape[1:])
        return torch.cat([randn_tensor(shape1, g, device, dtype) for g in generator], dim=0)
    if generator is not None and generator.device.type != device.type and generator.device.type == "cpu":
        return torch.randn(shape, generator=generator, device="cpu", dtype=dtype).to(device)
    if generator is not None and generator.device.type != device.type:
        raise ValueError(f"Cannot generate a {device} tensor from a generator of type {generator.device.type}.")
    return torch.randn(shape, generator=generator, device=device, dtype=dtype)


@dataclass
class WanPipelineOutput:
    frames: torch.Tensor


@dataclass
class CogVideoXPipelineOutput:
    frames: torch.Tensor


@dataclass
class HunyuanVideoPipelineOutput:
    frames: torch.Tensor


class PipelineCallback:  # diffusers.callbacks.PipelineCallback surface
    tensor_inputs: List[str] = []


class MultiPipelineCallbacks(PipelineCallback):
    pass


def load_image(image):
    from PIL import Image, ImageOps

    if isinstance(image, str):
        image = Image.open(image)
    image = ImageOps.exif_transpose(image)
    return image.convert("RGB")


def write_video(path, frames, fps):
    """run.py:121-133 replacement: frames = list of PIL images or an array [T, H, W, 3] in [0, 1] (or uint8) -> mp4.

    ``torchvision.io.write_video`` no longer exists in this image's torchvision and there is no ffmpeg binary / PyAV;
    OpenCV's FFMPEG-backed VideoWriter is what is available (mp4v fourcc)."""
    import cv2

    arr = np.stack([np.asarray(f) for f in frames]) if not isinstance(frames, np.ndarray) else frames
    if arr.dtype != np.uint8:
        arr = (np.clip(arr, 0, 1) * 255).round().astype(np.uint8)
    t, h, w, _ = arr.shape
    # run.py:127-133 asks for h264 (crf 18).  This image's OpenCV / FFmpeg build has no software H.264 encoder (only the
    # h264_v4l2m2m hardware wrapper, which needs a V4L2 device): try avc1 first, so a build that has libx264 / openh264 writes
    # what the reference writes, and fall back to MPEG-4 part 2 (mp4v) here.
    vw = None
    prev = os.environ.get("OPENCV_LOG_LEVEL")
    os.environ["OPENCV_LOG_LEVEL"] = "SILENT"
    try:
        for fourcc in ("avc1", "mp4v"):
            cand = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*fourcc), float(fps), (w, h))
            if cand.isOpened():
                vw = cand
                break
            cand.release()
    finally:
        if prev is None:
            os.environ.pop("OPENCV_LOG_LEVEL", None)
        else:
            os.environ["OPENCV_LOG_LEVEL"] = prev
    if vw is None:
        raise RuntimeError(f"cannot open a video writer for {path}")
    for f in arr:
        vw.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    vw.release()
    return path


class VideoProcessor:
    """preprocess: PIL / ndarray / tensor -> fp32 [B, 3, H, W] in [-1, 1]; postprocess_video: [B, C, T, H, W] -> np / pt / pil."""

    def __init__(self, vae_scale_factor: int = 8):
        self.vae_scale_factor = vae_scale_factor

    def preprocess(self, image, height=None, width=None):
        from PIL import Image

        if isinstance(image, Image.Image):
            image = [image]
        if isinstance(image, (list, tuple)) and isinstance(image[0], Image.Image):
            arrs = []
            for im in image:
                if height and width:
                    im = im.convert("RGB").resize((width, height), Image.LANCZOS)
                arrs.append(np.asarray(im.convert("RGB"), dtype=np.float32) / 255.0)
            t = torch.from_numpy(np.stack(arrs)).permute(0, 3, 1, 2)
        elif isinstance(image, np.ndarray):
            t = torch.from_numpy(image).float()
            t = t[None] if t.ndim == 3 else t
            t = t.permute(0, 3, 1, 2)
        elif torch.is_tensor(image):
            t = image.float()
            t = t[None] if t.ndim == 3 else t
        else:
            raise ValueError(f"unsupported image input {type(image)}")
        if height and width and tuple(t.shape[-2:]) != (height, width):
            t = F.interpolate(t, size=(height, width), mode="bilinear", align_corners=False)
        return 2.0 * t - 1.0

    def postprocess_video(self, video, output_type="np"):
        vids = []
        for v in video:  # [C, T, H, W]
            v = (v.float() / 2 + 0.5).clamp(0, 1).permute(1, 0, 2, 3)  # [T, C, H, W]
            if output_type == "pt":
                vids.append(v)
                continue
            arr = v.permute(0, 2, 3, 1).cpu().numpy()
            if output_type == "np":
                vids.append(arr)
            elif output_type == "pil":
                from PIL import Image

                vids.append([Image.fromarray((f * 255).round().astype("uint8")) for f in arr])
            else:
                raise ValueError(f"{output_type} does not exist. Please choose one of ['np', 'pt', 'pil']")
        if output_type == "np":
            return np.stack(vids)
        if output_type == "pt":
            return torch.stack(vids)
        return vids


class DiffusionPipelineBase:
    """register_modules / to / _execution_device / progress_bar / maybe_free_model_hooks of DiffusionPipeline."""

    def register_modules(self, **modules):
        self._modules = getattr(self, "_modules", {})
        for k, v in modules.items():
            self._modules[k] = v
            setattr(self, k, v)

    def to(self, device=None, dtype=None):
        self._device = torch.device(device) if device is not None else getattr(self, "_device", torch.device("cpu"))
        for k, m in self._modules.items():
            if m is not None and hasattr(m, "to") and not k.startswith("tokenizer") and k != "scheduler":
                r = m.to(self._device)
                if r is not None:
                    setattr(self, k, r)
                    self._modules[k] = r
        return self

    @property
    def device(self):
        return getattr(self, "_device", torch.device("cpu"))

    @property
    def _execution_device(self):
        return self.device

    @contextlib.contextmanager
    def progress_bar(self, total=None):
        try:
            from tqdm.auto import tqdm

            bar = tqdm(total=total, disable=getattr(self, "_progress_bar_disabled", False))
        except Exception:  # pragma: no cover
            bar = SimpleNamespace(update=lambda *a, **k: None, close=lambda: None)
        try:
            yield bar
        finally:
            bar.close()

    def set_progress_bar_config(self, disable=False, **kw):
        self._progress_bar_disabled = disable

    def maybe_free_model_hooks(self):
        pass


# --------------------------------------------------------------------------------------------------
# Synthetic stand-ins (once-per-video components; SURVEY 8(f) "next")
# --------------------------------------------------------------------------------------------------
class _Dist:
    def __init__(self, mean, logvar):
        self.mean, self.logvar = mean, logvar

    def mode(self):
        return self.mean

    def sample(self, generator=None):
        noise = randn_tensor(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + torch.exp(0.5 * self.logvar) * noise


class SyntheticVideoVAE(torch.nn.Module):
    """Latent-geometry-faithful placeholder for AutoencoderKL{Wan,CogVideoX,HunyuanVideo}: 8x spatial, 4x temporal
    (first frame kept), 3 <-> z_dim channels through a seeded 1x1x1 projection.  NOT the real VAE."""

    def __init__(self, z_dim=16, temporal=4, spatial=8, seed=0, scaling_factor=1.0, latents_mean=None, latents_std=None,
                 dtype=torch.float32):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.enc = torch.nn.Parameter(torch.randn(z_dim, 3, generator=g) * 0.5, requires_grad=False)
        self.dec = torch.nn.Parameter(torch.linalg.pinv(self.enc.data), requires_grad=False)
        self.temporal, self.spatial = temporal, spatial
        self.config = SimpleNamespace(
            z_dim=z_dim, latent_channels=z_dim, scaling_factor=scaling_factor, invert_scale_latents=False,
            scale_factor_temporal=temporal, scale_factor_spatial=spatial,
            latents_mean=latents_mean or [0.0] * z_dim, latents_std=latents_std or [1.0] * z_dim,
            block_out_channels=[1] * (int(math.log2(spatial)) + 1), temporal_compression_ratio=temporal,
            temperal_downsample=[False] * (int(math.log2(spatial)) - int(math.log2(temporal))) + [True] * int(math.log2(temporal)))
        self.to(dtype)

    @property
    def dtype(self):
        return self.enc.dtype

    def _pool_t(self, x):
        if x.shape[2] == 1:
            return x
        first, rest = x[:, :, :1], x[:, :, 1:]
        n = rest.shape[2] // self.temporal
        rest = rest[:, :, : n * self.temporal].unflatten(2, (n, self.temporal)).mean(3)
        return torch.cat([first, rest], dim=2)

    def encode(self, x):
        x = self._pool_t(x.to(self.dtype))
        B, C, T, H, W = x.shape
        x = F.avg_pool2d(x.transpose(1, 2).reshape(B * T, C, H, W), self.spatial).view(B, T, C, H // self.spatial, W // self.spatial)
        mean = torch.einsum("zc,btchw->bzthw", self.enc, x)
        return SimpleNamespace(latent_dist=_Dist(mean, torch.full_like(mean, -8.0)))

    def decode(self, z, return_dict=True):
        z = z.to(self.dtype)
        x = torch.einsum("cz,bzthw->bcthw", self.dec, z)
        B, C, T, H, W = x.shape
        x = F.interpolate(x.transpose(1, 2).reshape(B * T, C, H, W), scale_factor=self.spatial, mode="nearest")
        x = x.view(B, T, C, H * self.spatial, W * self.spatial).transpose(1, 2)
        if T > 1:
            x = torch.cat([x[:, :, :1], x[:, :, 1:].repeat_interleave(self.temporal, dim=2)], dim=2)
        x = x.clamp(-1, 1)
        return SimpleNamespace(sample=x) if return_dict else (x,)


class _Features(dict):
    """transformers.BatchFeature-like: attribute access + ``.to(device)``."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def to(self, device=None, dtype=None):
        return _Features({k: (v.to(device) if torch.is_tensor(v) else v) for k, v in self.items()})


class SyntheticTokenizer:
    """HF tokenizer call surface (wan:200-209, cog:244-253) without a vocabulary file: whitespace words -> CRC ids, one EOS
    (id 1), zero padding to ``max_length``.  Stand-in for the UMT5 / T5 sentencepiece tokenizers (no tokenizer files offline)."""

    def __init__(self, vocab_size: int = 256384):
        self.vocab_size = vocab_size

    def __call__(self, prompt, padding="max_length", max_length=512, truncation=True, add_special_tokens=True,
                 return_attention_mask=True, return_tensors="pt", **kw):
        import zlib

        prompt = [prompt] if isinstance(prompt, str) else list(prompt)
        ids = torch.zeros(len(prompt), max_length, dtype=torch.int64)
        mask = torch.zeros(len(prompt), max_length, dtype=torch.int64)
        for b, text in enumerate(prompt):
            toks = [2 + zlib.crc32(w.encode()) % (self.vocab_size - 2) for w in text.split()][: max_length - 1] + [1]
            ids[b, : len(toks)] = torch.tensor(toks)
            mask[b, : len(toks)] = 1
        return _Features(input_ids=ids, attention_mask=mask)


class SyntheticImageProcessor:
    """``CLIPImageProcessor`` call surface (wan:232): resize to 224 x 224 (bicubic), CLIP mean / std normalisation."""

    MEAN, STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)

    def __call__(self, images=None, return_tensors="pt", **kw):
        from PIL import Image

        images = images if isinstance(images, (list, tuple)) else [images]
        out = []
        for im in images:
            if isinstance(im, Image.Image):
                t = torch.from_numpy(np.asarray(im.convert("RGB").resize((224, 224), Image.BICUBIC), dtype=np.float32) / 255.0).permute(2, 0, 1)
            else:
                t = torch.as_tensor(im).float().cpu()
                t = t[0] if t.ndim == 4 else t
                t = (t + 1) / 2 if float(t.min()) < 0 else t
                t = F.interpolate(t[None], size=(224, 224), mode="bicubic", align_corners=False)[0].clamp(0, 1)
            out.append((t - torch.tensor(self.MEAN).view(3, 1, 1)) / torch.tensor(self.STD).view(3, 1, 1))
        return _Features(pixel_values=torch.stack(out))


class SyntheticTextEncoder:
    """Deterministic prompt -> [max_len, dim] embedding (hash-seeded); zero beyond the token count like Wan's
    zero-padding (wan:214-217).  NOT UMT5 / T5 / LLaVA."""

    def __init__(self, dim=4096, dtype=torch.bfloat16):
        self.dim, self.dtype = dim, dtype
        self.device = torch.device("cpu")

    def to(self, device=None, dtype=None):
        if device is not None:
            self.device = torch.device(device)
        return self

    def __call__(self, input_ids, attention_mask=None, **kw):
        """HF encoder call surface (wan:212, cog:258): ``last_hidden_state`` [B, L, dim], one seeded vector per token id."""
        ids = input_ids.cpu()
        uniq, inv = torch.unique(ids, return_inverse=True)
        table = torch.stack([torch.randn(self.dim, generator=torch.Generator().manual_seed(int(u) + 1)) for u in uniq])
        h = table[inv].to(self.device, self.dtype)
        return _EncOut(last_hidden_state=h)

    def embed(self, prompts: List[str], max_len: int, zero_pad=True) -> torch.Tensor:
        out = []
        for p in prompts:
            import zlib

            g = torch.Generator().manual_seed(zlib.crc32(p.encode()) & 0x7FFFFFFF)
            n = max(1, min(max_len, len(p.split()) + 2))
            e = torch.randn(max_len, self.dim, generator=g)
            if zero_pad:
                e[n:] = 0
            out.append(e)
        return torch.stack(out).to(self.device, self.dtype)


class _EncOut(SimpleNamespace):
    def __getitem__(self, i):
        return (self.last_hidden_state,)[i]


class SyntheticImageEncoder:
    """image -> [1, 257, 1280] penultimate-layer-like embedding.  NOT CLIP-ViT-H."""

    def __init__(self, tokens=257, dim=1280, dtype=torch.float32):
        self.tokens, self.dim, self.dtype = tokens, dim, dtype
        self.device = torch.device("cpu")

    def to(self, device=None, dtype=None):
        if device is not None:
            self.device = torch.device(device)
        return self

    def embed(self, image_tensor: torch.Tensor) -> torch.Tensor:
        seed = int(abs(float(image_tensor.float().mean())) * 1e6) % (2 ** 31 - 1)
        g = torch.Generator().manual_seed(seed)
        return torch.randn(1, self.tokens, self.dim, generator=g).to(self.device, self.dtype)

    def __call__(self, pixel_values=None, output_hidden_states=True, **kw):
        """``CLIPVisionModel`` call surface (wan:233-234): the pipeline reads ``hidden_states[-2]``."""
        h = torch.cat([self.embed(pv) for pv in pixel_values], dim=0)
        return SimpleNamespace(hidden_states=[h, h, h], last_hidden_state=h)
