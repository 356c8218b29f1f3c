"""HunyuanVideo image-to-video pipeline with Adaptive Low-pass Guidance -- B200-native drop-in.

Same module name, class name, constructor and ``__call__`` signature as the reference's
``pipeline_hunyuan_video_image2video_lowpass.HunyuanVideoImageToVideoPipeline`` (reference file:line cited per method);
the per-step work runs in ``libalg_b200.so``:

    low-pass filter of the first-frame latent   -> alg_lowpass_down_up / alg_lowpass_gaussian       (hy:1156-1167, 1220-1231)
    first-frame replacement + DiT forward (1-3 passes) -> alg_patch_gather + tcgen05 GEMM / attention (hy:1168-1252)
    (true-)CFG combine + FlowMatchEuler step on frames 1.. + re-prepend of the image frame -> alg_cfg_euler_step (hy:1254-1270)

Only what the reference can actually run is built (quirk q9): ``image_condition_type="token_replace"`` checkpoints and
in-latent filtering -- its pixel-space ``prepare_lp`` receives a PIL image and Wan-VAE config fields (hy:1166, 741-748)
and fails before reaching the filter.
"""
from __future__ import annotations

import inspect
import logging
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import numpy as np
import PIL.Image
import torch

import lp_utils
from alg_b200.hunyuan import HUNYUAN_VIDEO_I2V, HunyuanVideoTransformer3DModel
from alg_b200.pipeline_utils import (DiffusionPipelineBase, HunyuanVideoPipelineOutput, MultiPipelineCallbacks,
                                     PipelineCallback, SyntheticTextEncoder, SyntheticVideoVAE, VideoProcessor,
                                     randn_tensor)
from alg_b200.schedulers import FlowMatchEulerDiscreteScheduler

logger = logging.getLogger(__name__)

DEFAULT_PROMPT_TEMPLATE = {
    "template": (
        "<|start_header_id|>system<|end_header_id|>\n\n<image>\nDescribe the video by detailing the following aspects according to the reference image: "
        "1. The main content and theme of the video."
        "2. The color, shape, size, texture, quantity, text, and spatial relationships of the objects."
        "3. Actions, events, behaviors temporal relationships, physical movement changes of the objects."
        "4. background environment, light, style and atmosphere."
        "5. camera angles, movements, and transitions used in the video:<|eot_id|>\n\n"
        "<|start_header_id|>user<|end_header_id|>\n\n{}<|eot_id|>"
        "<|start_header_id|>assistant<|end_header_id|>\n\n"
    ),
    "crop_start": 103,
    "image_emb_start": 5,
    "image_emb_end": 581,
    "image_emb_len": 576,
    "double_return_token_id": 271,
}


def retrieve_timesteps(scheduler, num_inference_steps: Optional[int] = None, device=None,
                       timesteps: Optional[List[int]] = None, sigmas: Optional[List[float]] = None, **kwargs):
    """hy:152-208: calls ``scheduler.set_timesteps`` (custom ``timesteps`` / ``sigmas`` only if it accepts them)."""
    if timesteps is not None and sigmas is not None:
        raise ValueError("Only one of `timesteps` or `sigmas` can be passed. Please choose one to set custom values")
    if timesteps is not None:
        if "timesteps" not in set(inspect.signature(scheduler.set_timesteps).parameters.keys()):
            raise ValueError(f"The current scheduler class {scheduler.__class__}'s `set_timesteps` does not support custom"
                             f" timestep schedules. Please check whether you are using the correct scheduler.")
        scheduler.set_timesteps(timesteps=timesteps, device=device, **kwargs)
        timesteps = scheduler.timesteps
        num_inference_steps = len(timesteps)
    elif sigmas is not None:
        if "sigmas" not in set(inspect.signature(scheduler.set_timesteps).parameters.keys()):
            raise ValueError(f"The current scheduler class {scheduler.__class__}'s `set_timesteps` does not support custom"
                             f" sigmas schedules. Please check whether you are using the correct scheduler.")
        scheduler.set_timesteps(sigmas=sigmas, device=device, **kwargs)
        timesteps = scheduler.timesteps
        num_inference_steps = len(timesteps)
    else:
        scheduler.set_timesteps(num_inference_steps, device=device, **kwargs)
        timesteps = scheduler.timesteps
    return timesteps, num_inference_steps


def retrieve_latents(encoder_output, generator=None, sample_mode: str = "sample"):
    if hasattr(encoder_output, "latent_dist") and sample_mode == "sample":
        return encoder_output.latent_dist.sample(generator)
    if hasattr(encoder_output, "latent_dist") and sample_mode == "argmax":
        return encoder_output.latent_dist.mode()
    if hasattr(encoder_output, "latents"):
        return encoder_output.latents
    raise AttributeError("Could not access latents of provided encoder_output")


def _expand_input_ids_with_image_tokens(text_input_ids, prompt_attention_mask, max_sequence_length, image_token_index,
                                        image_emb_len, image_emb_start, image_emb_end, pad_token_id):
    """hy:107-149: every ``<image>`` token becomes ``image_emb_len`` slots; later tokens shift right; the attention mask is
    rebuilt from "not the pad token" and ``position_ids`` count the attended tokens (masked slots get position 1)."""
    is_image = text_input_ids == image_token_index
    n_image = torch.sum(is_image, dim=-1)
    rows, cols = torch.where(text_input_ids != image_token_index)
    expanded_len = max_sequence_length + (n_image.max() * (image_emb_len - 1))
    new_pos = torch.cumsum((is_image * (image_emb_len - 1) + 1), -1) - 1
    expanded_ids = torch.full((text_input_ids.shape[0], expanded_len), pad_token_id, dtype=text_input_ids.dtype,
                              device=text_input_ids.device)
    expanded_ids[rows, new_pos[rows, cols]] = text_input_ids[rows, cols]
    expanded_ids[rows, image_emb_start:image_emb_end] = image_token_index
    expanded_mask = torch.zeros((text_input_ids.shape[0], expanded_len), dtype=prompt_attention_mask.dtype,
                                device=prompt_attention_mask.device)
    arows, acols = torch.where(expanded_ids != pad_token_id)
    expanded_mask[arows, acols] = 1.0
    expanded_mask = expanded_mask.to(prompt_attention_mask.dtype)
    position_ids = (expanded_mask.cumsum(-1) - 1).masked_fill_((expanded_mask == 0), 1)
    return {"input_ids": expanded_ids, "attention_mask": expanded_mask, "position_ids": position_ids}


class HunyuanVideoImageToVideoPipeline(DiffusionPipelineBase):
    """Image-to-video generation with HunyuanVideo + ALG on the native sm_100a kernels (reference class: hy:224-280)."""

    model_cpu_offload_seq = "text_encoder->text_encoder_2->transformer->vae"
    _callback_tensor_inputs = ["latents", "prompt_embeds"]

    def __init__(self, text_encoder, tokenizer, transformer: HunyuanVideoTransformer3DModel, vae, scheduler, text_encoder_2,
                 tokenizer_2, image_processor):
        self.register_modules(vae=vae, text_encoder=text_encoder, tokenizer=tokenizer, transformer=transformer,
                              scheduler=scheduler, text_encoder_2=text_encoder_2, tokenizer_2=tokenizer_2,
                              image_processor=image_processor)
        has_vae = getattr(self, "vae", None) is not None
        self.vae_scaling_factor = self.vae.config.scaling_factor if has_vae else 0.476986
        self.vae_scale_factor_temporal = getattr(self.vae, "temporal_compression_ratio", 4) if has_vae else 4
        self.vae_scale_factor_spatial = getattr(self.vae, "spatial_compression_ratio", 8) if has_vae else 8
        self.video_processor = VideoProcessor(vae_scale_factor=self.vae_scale_factor_spatial)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, transformer=None, vae=None, torch_dtype=torch.float16,
                        cache_dir=None, synthetic: Optional[bool] = None, allow_synthetic_aux: bool = False,
                        seed: int = 0, device="cuda", tokenizer=None, text_encoder=None, tokenizer_2=None, text_encoder_2=None,
                        image_processor=None, **config_overrides):
        """run.py:71-81.  A local diffusers snapshot loads the real DiT weights and scheduler config plus the native LLaVA /
        CLIP-L prompt encoders from its ``text_encoder`` / ``text_encoder_2`` folders; offline there are no checkpoints:
        ``synthetic=True`` (or ``ALG_SYNTHETIC=1``) builds the true HunyuanVideo-I2V architecture with seeded random weights
        directly on ``device``."""
        import os

        if synthetic is None:
            synthetic = os.environ.get("ALG_SYNTHETIC", "0") == "1" or str(pretrained_model_name_or_path).startswith("synthetic")
        scheduler = None
        if not synthetic:
            from alg_b200 import checkpoint, encoders, llava

            snap = checkpoint.resolve_snapshot(str(pretrained_model_name_or_path), cache_dir)
            if snap is None:
                raise FileNotFoundError(
                    f"no local diffusers snapshot for {pretrained_model_name_or_path!r}: there is no network, so pass a "
                    "directory (or a hub id already present under cache_dir), or synthetic=True / ALG_SYNTHETIC=1")
            if transformer is None:
                transformer, scheduler = checkpoint.build_from_snapshot(HunyuanVideoTransformer3DModel, FlowMatchEulerDiscreteScheduler, snap, device)
            else:
                scheduler = FlowMatchEulerDiscreteScheduler.from_config(checkpoint.scheduler_config(snap))
            if text_encoder is None:  # native LLaVA-Llama-3 (fp32 arithmetic, reported dtype = torch_dtype) + the snapshot's tokenizer
                tokenizer, text_encoder = checkpoint.load_text_stack(snap, llava.LlavaForConditionalGeneration, device, tokenizer,
                                                                     torch_dtype=torch_dtype)
            if text_encoder_2 is None:  # native CLIP-L text tower
                tokenizer_2, text_encoder_2 = checkpoint.load_text_stack(snap, encoders.CLIPTextModel, device, tokenizer_2,
                                                                         tokenizer_dir="tokenizer_2", encoder_dir="text_encoder_2")
            if image_processor is None and os.path.isdir(os.path.join(snap, "image_processor")):
                from transformers import CLIPImageProcessor

                image_processor = CLIPImageProcessor.from_pretrained(os.path.join(snap, "image_processor"))
            if vae is None and os.path.isdir(os.path.join(snap, "vae")):  # native AutoencoderKLHunyuanVideo (float32 arithmetic)
                from alg_b200.vae_hunyuan import AutoencoderKLHunyuanVideo

                vae = AutoencoderKLHunyuanVideo.from_pretrained(snap, device=device)
            if (vae is None or text_encoder is None or text_encoder_2 is None or image_processor is None) and not allow_synthetic_aux:
                raise NotImplementedError(checkpoint.AUX_MESSAGE)
        if transformer is None:
            transformer = HunyuanVideoTransformer3DModel.from_synthetic(seed=seed, device=device, **config_overrides)
        if synthetic and os.environ.get("ALG_NATIVE_ENCODERS", "0") == "1":
            # synthetic weights at the TRUE encoder architectures (LLaVA-Llama-3-8B, CLIP-L) through the native kernels; the
            # tokenizers have no offline vocabulary, so prompts still need tokenizer / tokenizer_2 objects from the caller
            from alg_b200 import encoders, llava

            text_encoder = text_encoder or llava.LlavaForConditionalGeneration.from_synthetic(seed=seed, device=device, torch_dtype=torch_dtype)
            text_encoder_2 = text_encoder_2 or encoders.CLIPTextModel.from_synthetic(seed=seed, device=device, **encoders.CLIP_L_TEXT)
        if vae is None and synthetic and os.environ.get("ALG_NATIVE_VAE", "0") == "1":
            from alg_b200.vae_hunyuan import AutoencoderKLHunyuanVideo  # seeded weights at the true architecture, native kernels

            vae = AutoencoderKLHunyuanVideo.from_synthetic(seed=seed, device=device)
        if vae is None:
            vae = SyntheticVideoVAE(z_dim=transformer.config.in_channels, scaling_factor=0.476986, dtype=torch_dtype)
            vae.temporal_compression_ratio, vae.spatial_compression_ratio = 4, 8
        return cls(text_encoder=text_encoder or SyntheticTextEncoder(transformer.config.text_embed_dim, torch_dtype), tokenizer=tokenizer,
                   transformer=transformer, vae=vae, scheduler=scheduler or FlowMatchEulerDiscreteScheduler(shift=7.0),
                   text_encoder_2=text_encoder_2 or SyntheticTextEncoder(transformer.config.pooled_projection_dim, torch_dtype),
                   tokenizer_2=tokenizer_2, image_processor=image_processor)

    # ------------------------------------------------------------------------------------------------
    # once-per-video conditioning (hy:282-492).  tokenizer / text_encoder (LLaVA-Llama-3) / image_processor and tokenizer_2 /
    # text_encoder_2 (CLIP-L) are driven through the transformers call surface the reference uses, so real HF objects and the
    # native encoders (alg_b200/llava.py, alg_b200/encoders.py::CLIPTextModel) are interchangeable; the hash-seeded
    # SyntheticTextEncoder stand-ins (synthetic weights only) keep their own short path and produce the same shapes.
    def _get_llama_prompt_embeds(self, image, prompt, prompt_template, num_videos_per_prompt=1, device=None, dtype=None,
                                 max_sequence_length: int = 256, num_hidden_layers_to_skip: int = 2,
                                 image_embed_interleave: int = 2):
        device = device or self._execution_device
        dtype = dtype or self.text_encoder.dtype
        prompt = [prompt] if isinstance(prompt, str) else prompt
        if isinstance(self.text_encoder, SyntheticTextEncoder):
            n_img = prompt_template.get("image_emb_len", 576) // max(image_embed_interleave, 1)
            n_img = min(n_img, 16)  # the synthetic encoder keeps the image-token block short
            total = n_img + max_sequence_length
            embeds = self.text_encoder.embed(prompt, total, zero_pad=False).to(device=device, dtype=dtype)
            mask = torch.zeros(len(prompt), total, device=device, dtype=torch.long)
            for b, p in enumerate(prompt):
                mask[b, : n_img + max(1, min(max_sequence_length, len(p.split()) + 2))] = 1
            return embeds.repeat_interleave(num_videos_per_prompt, dim=0), mask.repeat_interleave(num_videos_per_prompt, dim=0)

        # hy:297-337: template -> ids -> <image> expanded to image_emb_len slots -> LLaVA hidden state (skip + 1) from the end
        prompt = [prompt_template["template"].format(p) for p in prompt]
        crop_start = prompt_template.get("crop_start", None)
        image_emb_len = prompt_template.get("image_emb_len", 576)
        image_emb_start = prompt_template.get("image_emb_start", 5)
        image_emb_end = prompt_template.get("image_emb_end", 581)
        double_return_token_id = prompt_template.get("double_return_token_id", 271)
        if crop_start is None:
            template_ids = self.tokenizer(prompt_template["template"], padding="max_length", return_tensors="pt",
                                          return_length=False, return_overflowing_tokens=False, return_attention_mask=False)
            # minus <|start_header_id|>, <|end_header_id|>, assistant, <|eot_id|> and the {} placeholder (hy:308-310)
            crop_start = template_ids["input_ids"].shape[-1] - 5
        max_sequence_length += crop_start
        text_inputs = self.tokenizer(prompt, max_length=max_sequence_length, padding="max_length", truncation=True,
                                     return_tensors="pt", return_length=False, return_overflowing_tokens=False,
                                     return_attention_mask=True)
        text_input_ids = text_inputs.input_ids.to(device=device)
        prompt_attention_mask = text_inputs.attention_mask.to(device=device)
        pixel_values = self.image_processor(image, return_tensors="pt").pixel_values.to(device)
        expanded = _expand_input_ids_with_image_tokens(
            text_input_ids, prompt_attention_mask, max_sequence_length, self.text_encoder.config.image_token_index,
            image_emb_len, image_emb_start, image_emb_end, self.text_encoder.config.pad_token_id)
        prompt_embeds = self.text_encoder(**expanded, pixel_values=pixel_values,
                                          output_hidden_states=True).hidden_states[-(num_hidden_layers_to_skip + 1)]
        prompt_embeds = prompt_embeds.to(dtype=dtype)

        if crop_start is not None and crop_start > 0:
            # hy:342-399: drop the system template and the 4-token assistant header; image slots become their own block in front
            text_crop_start = crop_start - 1 + image_emb_len
            rows, cols = torch.where(text_input_ids == double_return_token_id)
            if cols.shape[0] == 3:  # the prompt was truncated before the assistant header's "\n\n" (hy:346-351)
                cols = torch.cat((cols, torch.tensor([text_input_ids.shape[-1]], device=cols.device)))
                rows = torch.cat((rows, torch.tensor([0], device=rows.device)))
            last_dr = cols.reshape(text_input_ids.shape[0], -1)[:, -1]
            text_list, mask_list, image_list, image_mask_list = [], [], [], []
            for i in range(text_input_ids.shape[0]):
                dr = int(last_dr[i].item())
                a0, a1 = dr - 1 + image_emb_len - 4, dr - 1 + image_emb_len  # the assistant header inside the expanded sequence
                text_list.append(torch.cat([prompt_embeds[i, text_crop_start:a0], prompt_embeds[i, a1:]]))
                mask_list.append(torch.cat([prompt_attention_mask[i, crop_start:dr - 4], prompt_attention_mask[i, dr:]]))
                image_list.append(prompt_embeds[i, image_emb_start:image_emb_end])
                image_mask_list.append(torch.ones(image_list[-1].shape[0]).to(prompt_embeds.device).to(prompt_attention_mask.dtype))
            text_embeds, text_mask = torch.stack(text_list), torch.stack(mask_list)
            image_embeds, image_mask = torch.stack(image_list), torch.stack(image_mask_list)
            if 0 < image_embed_interleave < 6:
                image_embeds = image_embeds[:, ::image_embed_interleave, :]
                image_mask = image_mask[:, ::image_embed_interleave]
            assert text_embeds.shape[0] == text_mask.shape[0] and image_embeds.shape[0] == image_mask.shape[0]
            prompt_embeds = torch.cat([image_embeds, text_embeds], dim=1)
            prompt_attention_mask = torch.cat([image_mask, text_mask], dim=1)
        return prompt_embeds, prompt_attention_mask

    def _get_clip_prompt_embeds(self, prompt, num_videos_per_prompt=1, device=None, dtype=None, max_sequence_length: int = 77):
        device = device or self._execution_device
        dtype = dtype or self.text_encoder_2.dtype
        prompt = [prompt] if isinstance(prompt, str) else prompt
        if isinstance(self.text_encoder_2, SyntheticTextEncoder):
            pooled = self.text_encoder_2.embed(prompt, 1, zero_pad=False)[:, 0].to(device=device, dtype=dtype)
            return pooled.repeat_interleave(num_videos_per_prompt, dim=0)
        # hy:421-452: ids only (no attention mask), pooled = the end-of-text token's final state; returned as the encoder made it
        text_inputs = self.tokenizer_2(prompt, padding="max_length", max_length=max_sequence_length, truncation=True,
                                       return_tensors="pt")
        text_input_ids = text_inputs.input_ids
        untruncated_ids = self.tokenizer_2(prompt, padding="longest", return_tensors="pt").input_ids
        if untruncated_ids.shape[-1] >= text_input_ids.shape[-1] and not torch.equal(text_input_ids, untruncated_ids):
            removed_text = self.tokenizer_2.batch_decode(untruncated_ids[:, max_sequence_length - 1: -1])
            logger.warning("The following part of your input was truncated because CLIP can only handle sequences up to"
                           f" {max_sequence_length} tokens: {removed_text}")
        return self.text_encoder_2(text_input_ids.to(device), output_hidden_states=False).pooler_output

    def encode_prompt(self, image, prompt, prompt_2=None, prompt_template: Dict[str, Any] = DEFAULT_PROMPT_TEMPLATE,
                      num_videos_per_prompt: int = 1, prompt_embeds=None, pooled_prompt_embeds=None,
                      prompt_attention_mask=None, device=None, dtype=None, max_sequence_length: int = 256,
                      image_embed_interleave: int = 2) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        if prompt_embeds is None:
            prompt_embeds, prompt_attention_mask = self._get_llama_prompt_embeds(
                image, prompt, prompt_template, num_videos_per_prompt, device=device, dtype=dtype,
                max_sequence_length=max_sequence_length, image_embed_interleave=image_embed_interleave)
        if pooled_prompt_embeds is None:
            if prompt_2 is None:
                prompt_2 = prompt
            # the reference passes `prompt`, not `prompt_2`, to CLIP (hy:484-486, quirk q9)
            pooled_prompt_embeds = self._get_clip_prompt_embeds(prompt, num_videos_per_prompt, device=device, dtype=dtype,
                                                                max_sequence_length=77)
        return prompt_embeds, pooled_prompt_embeds, prompt_attention_mask

    def check_inputs(self, prompt, prompt_2, height, width, prompt_embeds=None, callback_on_step_end_tensor_inputs=None,
                     prompt_template=None, true_cfg_scale=1.0, guidance_scale=1.0):
        """Same ValueErrors, in the same order, as hy:494-548."""
        if height % 16 != 0 or width % 16 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 16 but are {height} and {width}.")
        if callback_on_step_end_tensor_inputs is not None and not all(
                k in self._callback_tensor_inputs for k in callback_on_step_end_tensor_inputs):
            raise ValueError(f"`callback_on_step_end_tensor_inputs` has to be in {self._callback_tensor_inputs}, but found "
                             f"{[k for k in callback_on_step_end_tensor_inputs if k not in self._callback_tensor_inputs]}")
        if prompt is not None and prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `prompt`: {prompt} and `prompt_embeds`: {prompt_embeds}. Please make sure to"
                             " only forward one of the two.")
        elif prompt_2 is not None and prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `prompt_2`: {prompt_2} and `prompt_embeds`: {prompt_embeds}. Please make sure to"
                             " only forward one of the two.")
        elif prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`. Cannot leave both `prompt` and `prompt_embeds` undefined.")
        elif prompt is not None and (not isinstance(prompt, str) and not isinstance(prompt, list)):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")
        elif prompt_2 is not None and (not isinstance(prompt_2, str) and not isinstance(prompt_2, list)):
            raise ValueError(f"`prompt_2` has to be of type `str` or `list` but is {type(prompt_2)}")
        if prompt_template is not None:
            if not isinstance(prompt_template, dict):
                raise ValueError(f"`prompt_template` has to be of type `dict` but is {type(prompt_template)}")
            if "template" not in prompt_template:
                raise ValueError(f"`prompt_template` has to contain a key `template` but only found {prompt_template.keys()}")
        if true_cfg_scale > 1.0 and guidance_scale > 1.0:
            logger.warning("Both `true_cfg_scale` and `guidance_scale` are greater than 1.0. This will result in both "
                           "classifier-free guidance and embedded-guidance to be applied. This is not recommended "
                           "as it may lead to higher memory usage, slower inference and potentially worse results.")

    def prepare_latents(self, image: torch.Tensor, batch_size: int, num_channels_latents: int = 32, height: int = 720,
                        width: int = 1280, num_frames: int = 129, dtype=None, device=None, generator=None, latents=None,
                        image_condition_type: str = "latent_concat", i2v_stable: bool = False):
        """Initial noise + VAE latent of the image (hy:550-599); token_replace keeps the single first latent frame."""
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an effective batch"
                             f" size of {batch_size}. Make sure the batch size matches the length of the generators.")
        num_latent_frames = (num_frames - 1) // self.vae_scale_factor_temporal + 1
        latent_height, latent_width = height // self.vae_scale_factor_spatial, width // self.vae_scale_factor_spatial
        shape = (batch_size, num_channels_latents, num_latent_frames, latent_height, latent_width)
        image = image.unsqueeze(2)  # [B, C, 1, H, W]
        if isinstance(generator, list):
            image_latents = [retrieve_latents(self.vae.encode(image[i].unsqueeze(0)), generator[i], "argmax") for i in range(batch_size)]
        else:
            image_latents = [retrieve_latents(self.vae.encode(img.unsqueeze(0)), generator, "argmax") for img in image]
        image_latents = torch.cat(image_latents, dim=0).to(dtype) * self.vae_scaling_factor
        if latents is None:
            latents = randn_tensor(shape, generator=generator, device=device, dtype=dtype)
        else:
            latents = latents.to(device=device, dtype=dtype)
        if i2v_stable:
            image_latents = image_latents.repeat(1, 1, num_latent_frames, 1, 1)
            t = torch.tensor([0.999]).to(device=device)
            latents = latents * t + image_latents * (1 - t)
        if image_condition_type == "token_replace":
            image_latents = image_latents[:, :, :1]
        return latents, image_latents

    def enable_vae_slicing(self):
        getattr(self.vae, "enable_slicing", lambda: None)()

    def disable_vae_slicing(self):
        getattr(self.vae, "disable_slicing", lambda: None)()

    def enable_vae_tiling(self):
        getattr(self.vae, "enable_tiling", lambda: None)()

    def disable_vae_tiling(self):
        getattr(self.vae, "disable_tiling", lambda: None)()

    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def num_timesteps(self):
        return self._num_timesteps

    @property
    def attention_kwargs(self):
        return self._attention_kwargs

    @property
    def current_timestep(self):
        return self._current_timestep

    @property
    def interrupt(self):
        return self._interrupt

    def prepare_lp(self, lp_filter_type, lp_blur_sigma, lp_blur_kernel_size, lp_resize_factor, generator, num_frames,
                   use_low_pass_guidance, lp_filter_in_latent, orig_image_latents, orig_image_tensor, last_image=None):
        """Low-passed first-frame latent (hy:650-792), in-latent mode: one CUDA launch on [1, 16, 1, h, w]."""
        if not use_low_pass_guidance:
            return None
        if not lp_filter_in_latent:
            # hy:1166 hands the PIL image to this path and hy:741-748 read Wan-VAE config fields: it cannot run (quirk q9)
            raise NotImplementedError("HunyuanVideo ALG filters in latent space only (the reference's pixel-space branch "
                                      "fails on the PIL image it is given); set lp_filter_in_latent=True")
        lp = lp_utils.apply_low_pass_filter(orig_image_latents, filter_type=lp_filter_type, blur_sigma=lp_blur_sigma,
                                            blur_kernel_size=lp_blur_kernel_size, resize_factor=lp_resize_factor)
        # hy:781-787 tests size(1) (16 channels) against patch_size 2: never prepends (quirk q7)
        assert lp.size(1) % self.transformer.config.patch_size == 0
        return lp.to(dtype=orig_image_latents.dtype)

    # ------------------------------------------------------------------------------------------------
    def denoise_step(self, i, t, latents, image_latents, pos, neg, guidance, num_frames, num_inference_steps,
                     alg: Dict[str, Any], true_cfg_scale: float, lp_on_noisy_latent: bool = False, generator=None):
        """One iteration of hy:1126-1270 for a single sample.  ``t`` is the scheduler's fp32 timestep (0-dim tensor or
        float); ``pos`` / ``neg`` = (prompt_embeds [1, L, D], pooled [1, P], mask [1, L]) (``neg`` None: no true CFG).
        Returns (new fp32 latents [1, 16, T, H, W], bf16 noise prediction of all passes)."""
        do_true_cfg = true_cfg_scale > 1 and neg is not None
        use_lp = alg["use_low_pass_guidance"]
        tdt = self.transformer.dtype

        def strength_and_lp():
            s = lp_utils.get_lp_strength(
                step_index=i, total_steps=num_inference_steps, lp_strength_schedule_type=alg["lp_strength_schedule_type"],
                schedule_interval_start_time=alg["schedule_interval_start_time"],
                schedule_interval_end_time=alg["schedule_interval_end_time"],
                schedule_linear_start_weight=alg["schedule_linear_start_weight"],
                schedule_linear_end_weight=alg["schedule_linear_end_weight"],
                schedule_linear_end_time=alg["schedule_linear_end_time"], schedule_exp_decay_rate=alg["schedule_exp_decay_rate"])
            sigma = alg["lp_blur_sigma"] * s
            ksize = alg["lp_blur_kernel_size"] * s if alg["schedule_blur_kernel_size"] else alg["lp_blur_kernel_size"]
            factor = 1.0 - (1.0 - alg["lp_resize_factor"]) * s
            if alg.get("enable_lp_img_embeds", False):
                assert False, "Low-pass filter on image embeds is not supported in HunyuanVideo pipeline. Please set enable_lp_img_embeds = False"
            lp = self.prepare_lp(lp_filter_type=alg["lp_filter_type"], lp_blur_sigma=sigma, lp_blur_kernel_size=ksize,
                                 lp_resize_factor=factor, generator=generator, num_frames=num_frames,
                                 use_low_pass_guidance=True, lp_filter_in_latent=alg["lp_filter_in_latent"],
                                 orig_image_latents=image_latents, orig_image_tensor=None)
            return s, lp

        if do_true_cfg and use_lp:
            s, lp = strength_and_lp()
            if s == 0.0 or lp_on_noisy_latent:  # hy:1168: the LP latent is computed, then unused
                firsts, ctx = [image_latents, image_latents], [neg, pos]
            else:
                firsts, ctx = [image_latents, lp, lp], [neg, neg, pos]
        elif do_true_cfg:
            firsts, ctx = [image_latents, image_latents], [neg, pos]
        elif not use_lp:
            firsts, ctx = [image_latents], [pos]
        else:  # ALG without true CFG (the shipped yaml): single pass on the low-passed first frame (hy:1196-1235)
            s, lp = strength_and_lp()
            firsts, ctx = [lp], [pos]

        t_model = float(torch.as_tensor(t, dtype=torch.float32).to(tdt))  # hy:1237: the timestep is cast to bf16
        T, H, W = latents.shape[2:]
        noise_pred = torch.empty(len(firsts), latents.shape[1], T, H, W, device=latents.device, dtype=tdt)
        for p, (first, (emb, pooled, mask)) in enumerate(zip(firsts, ctx)):
            n_valid = int(mask[0].sum().item()) if torch.is_tensor(mask) else int(mask)
            self.transformer.forward_pass(latents[0], first[0], emb[0, :n_valid], pooled[0], t_model, guidance, out=noise_pred[p])
        latents = self.scheduler.step_cfg_frames(noise_pred, true_cfg_scale, latents, image_latents)
        return latents, noise_pred

    @torch.no_grad()
    def __call__(
        self,
        image: PIL.Image.Image,
        prompt: Union[str, List[str]] = None,
        prompt_2: Union[str, List[str]] = None,
        negative_prompt: Union[str, List[str]] = "bad quality",
        negative_prompt_2: Union[str, List[str]] = None,
        height: int = 720,
        width: int = 1280,
        num_frames: int = 129,
        num_inference_steps: int = 50,
        sigmas: List[float] = None,
        true_cfg_scale: float = 1.0,
        guidance_scale: float = 1.0,
        num_videos_per_prompt: Optional[int] = 1,
        generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
        latents: Optional[torch.Tensor] = None,
        prompt_embeds: Optional[torch.Tensor] = None,
        pooled_prompt_embeds: Optional[torch.Tensor] = None,
        prompt_attention_mask: Optional[torch.Tensor] = None,
        negative_prompt_embeds: Optional[torch.Tensor] = None,
        negative_pooled_prompt_embeds: Optional[torch.Tensor] = None,
        negative_prompt_attention_mask: Optional[torch.Tensor] = None,
        output_type: Optional[str] = "pil",
        return_dict: bool = True,
        attention_kwargs: Optional[Dict[str, Any]] = None,
        callback_on_step_end: Optional[
            Union[Callable[[int, int, Dict], None], PipelineCallback, MultiPipelineCallbacks]
        ] = None,
        callback_on_step_end_tensor_inputs: List[str] = ["latents"],
        prompt_template: Dict[str, Any] = DEFAULT_PROMPT_TEMPLATE,
        max_sequence_length: int = 256,
        image_embed_interleave: Optional[int] = None,
        use_low_pass_guidance: bool = False,
        lp_filter_type: str = "none",
        lp_filter_in_latent: bool = False,
        lp_blur_sigma: float = 15.0,
        lp_blur_kernel_size: float = 0.02734375,
        lp_resize_factor: float = 0.25,
        lp_strength_schedule_type: str = "none",
        schedule_blur_kernel_size: bool = False,
        schedule_interval_start_time: float = 0.0,
        schedule_interval_end_time: float = 0.05,
        schedule_linear_start_weight: float = 1.0,
        schedule_linear_end_weight: float = 0.0,
        schedule_linear_end_time: float = 0.5,
        schedule_exp_decay_rate: float = 10.0,
        lp_on_noisy_latent=False,
        enable_lp_img_embeds=False,
        i2v_stable=False,
    ):
        """Generate a video (hy:796-1308).  Arguments, defaults and return type are those of the reference."""
        if isinstance(callback_on_step_end, (PipelineCallback, MultiPipelineCallbacks)):
            callback_on_step_end_tensor_inputs = callback_on_step_end.tensor_inputs
        self.check_inputs(prompt, prompt_2, height, width, prompt_embeds, callback_on_step_end_tensor_inputs, prompt_template,
                          true_cfg_scale, guidance_scale)
        cfg = self.transformer.config
        image_condition_type = cfg.image_condition_type
        has_neg_prompt = negative_prompt is not None or (negative_prompt_embeds is not None and negative_pooled_prompt_embeds is not None)
        do_true_cfg = true_cfg_scale > 1 and has_neg_prompt
        image_embed_interleave = (image_embed_interleave if image_embed_interleave is not None
                                  else (2 if image_condition_type == "latent_concat" else 4 if image_condition_type == "token_replace" else 1))
        self._guidance_scale = guidance_scale
        self._attention_kwargs = attention_kwargs
        self._current_timestep = None
        self._interrupt = False
        device = self._execution_device

        if prompt is not None and isinstance(prompt, str):
            batch_size = 1
        elif prompt is not None and isinstance(prompt, list):
            batch_size = len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]
        if batch_size * num_videos_per_prompt != 1:
            raise NotImplementedError("the native loop runs one sample per GPU (independent samples shard across GPUs)")

        vae_dtype = self.vae.dtype
        image_tensor = self.video_processor.preprocess(image, height, width).to(device, vae_dtype)
        num_channels_latents = cfg.in_channels  # token_replace (hy:1021-1024)
        latents, image_latents = self.prepare_latents(image_tensor, batch_size * num_videos_per_prompt, num_channels_latents,
                                                      height, width, num_frames, torch.float32, device, generator, latents,
                                                      image_condition_type, i2v_stable)

        transformer_dtype = self.transformer.dtype
        prompt_embeds, pooled_prompt_embeds, prompt_attention_mask = self.encode_prompt(
            image=image, prompt=prompt, prompt_2=prompt_2, prompt_template=prompt_template,
            num_videos_per_prompt=num_videos_per_prompt, prompt_embeds=prompt_embeds, pooled_prompt_embeds=pooled_prompt_embeds,
            prompt_attention_mask=prompt_attention_mask, device=device, max_sequence_length=max_sequence_length,
            image_embed_interleave=image_embed_interleave)
        prompt_embeds = prompt_embeds.to(device, transformer_dtype).contiguous()
        prompt_attention_mask = prompt_attention_mask.to(device)
        pooled_prompt_embeds = pooled_prompt_embeds.to(device, transformer_dtype).contiguous()
        neg = None
        if do_true_cfg:
            black_image = PIL.Image.new("RGB", (width, height), 0)
            negative_prompt_embeds, negative_pooled_prompt_embeds, negative_prompt_attention_mask = self.encode_prompt(
                image=black_image, prompt=negative_prompt, prompt_2=negative_prompt_2, prompt_template=prompt_template,
                num_videos_per_prompt=num_videos_per_prompt, prompt_embeds=negative_prompt_embeds,
                pooled_prompt_embeds=negative_pooled_prompt_embeds, prompt_attention_mask=negative_prompt_attention_mask,
                device=device, max_sequence_length=max_sequence_length, image_embed_interleave=image_embed_interleave)
            neg = (negative_prompt_embeds.to(device, transformer_dtype).contiguous(),
                   negative_pooled_prompt_embeds.to(device, transformer_dtype).contiguous(),
                   negative_prompt_attention_mask.to(device))

        sigmas = np.linspace(1.0, 0.0, num_inference_steps + 1)[:-1] if sigmas is None else sigmas
        timesteps, num_inference_steps = retrieve_timesteps(self.scheduler, num_inference_steps, device, sigmas=sigmas)

        guidance = None
        if cfg.guidance_embeds:  # hy:1115-1119: bf16(guidance_scale) * 1000 (6.0 -> 6016 in bf16)
            guidance = float((torch.tensor([guidance_scale] * latents.shape[0], dtype=transformer_dtype, device=device) * 1000.0)[0])

        alg = dict(use_low_pass_guidance=use_low_pass_guidance, lp_filter_type=lp_filter_type,
                   lp_filter_in_latent=lp_filter_in_latent, lp_blur_sigma=lp_blur_sigma,
                   lp_blur_kernel_size=lp_blur_kernel_size, lp_resize_factor=lp_resize_factor,
                   lp_strength_schedule_type=lp_strength_schedule_type, schedule_blur_kernel_size=schedule_blur_kernel_size,
                   schedule_interval_start_time=schedule_interval_start_time,
                   schedule_interval_end_time=schedule_interval_end_time,
                   schedule_linear_start_weight=schedule_linear_start_weight,
                   schedule_linear_end_weight=schedule_linear_end_weight, schedule_linear_end_time=schedule_linear_end_time,
                   schedule_exp_decay_rate=schedule_exp_decay_rate, enable_lp_img_embeds=enable_lp_img_embeds)

        num_warmup_steps = len(timesteps) - num_inference_steps * self.scheduler.order
        self._num_timesteps = len(timesteps)
        timesteps_host = timesteps.float().cpu()
        # the key-padding masks are step-invariant: read the valid-prefix lengths once, not once per step
        n_valid_pos = int(prompt_attention_mask[0].sum().item())
        if neg is not None:
            neg = (neg[0], neg[1], int(neg[2][0].sum().item()))
        with self.progress_bar(total=num_inference_steps) as progress_bar:
            for i in range(len(timesteps)):
                if self.interrupt:
                    continue
                t = timesteps[i]
                self._current_timestep = t
                pos = (prompt_embeds, pooled_prompt_embeds, n_valid_pos)
                latents, _ = self.denoise_step(i, timesteps_host[i], latents, image_latents, pos, neg, guidance, num_frames,
                                               num_inference_steps, alg, true_cfg_scale, lp_on_noisy_latent, generator)
                if callback_on_step_end is not None:
                    scope = dict(latents=latents, prompt_embeds=prompt_embeds)
                    outputs = callback_on_step_end(self, i, t, {k: scope[k] for k in callback_on_step_end_tensor_inputs})
                    latents = outputs.pop("latents", latents)
                    prompt_embeds = outputs.pop("prompt_embeds", prompt_embeds)
                if i == len(timesteps) - 1 or ((i + 1) > num_warmup_steps and (i + 1) % self.scheduler.order == 0):
                    progress_bar.update()
        self._current_timestep = None

        if not output_type == "latent":
            z = latents.to(self.vae.dtype) / self.vae_scaling_factor
            video = self.vae.decode(z, return_dict=False)[0]
            video = self.video_processor.postprocess_video(video, output_type=output_type)
        else:
            video = latents
        self.maybe_free_model_hooks()
        if not return_dict:
            return (video,)
        return HunyuanVideoPipelineOutput(frames=video)
