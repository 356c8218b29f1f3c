"""CogVideoX image-to-video pipeline with Adaptive Low-pass Guidance -- B200-native drop-in.

Same module name, class name, constructor and ``__call__`` signature as the reference's
``pipeline_cogvideox_image2video_lowpass.CogVideoXImageToVideoPipeline`` (reference file:line cited per method); the
per-step work runs in ``libalg_b200.so``:

    low-pass filter of the image condition       -> alg_lowpass_gaussian / alg_lowpass_down_up       (cog:1043-1057)
    model-input assembly + DiT forward (2-3 passes) -> alg_patch_gather + tcgen05 GEMM / attention     (cog:1059-1090)
    fp32 CFG combine + DDIM (v-prediction) step + cast back to bf16 -> alg_cfg_ddim_step             (cog:1091-1123)

Latent layout is the reference's ``[B, F, C, H, W]`` in ``prompt_embeds.dtype`` (bf16).  The ``[latents]*3`` /
``cat(dim=2)`` model input is never materialised: the patch gather reads ``latents`` and ``image_latents`` /
``lp_image_latents`` in place.
"""
from __future__ import annotations

import inspect
import math
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import PIL.Image
import torch

import lp_utils
from alg_b200.cogvideox import COGVIDEOX_5B_I2V, CogVideoXTransformer3DModel
from alg_b200.embeddings import get_3d_rotary_pos_embed, get_resize_crop_region_for_grid
from alg_b200.pipeline_utils import (CogVideoXPipelineOutput, DiffusionPipelineBase, MultiPipelineCallbacks,
                                     PipelineCallback, SyntheticTextEncoder, SyntheticTokenizer, SyntheticVideoVAE, VideoProcessor,
                                     randn_tensor)
from alg_b200.schedulers import CogVideoXDDIMScheduler, CogVideoXDPMScheduler
from alg_b200.vae_cogvideox import AutoencoderKLCogVideoX

PipelineImageInput = Union[PIL.Image.Image, torch.Tensor, List[PIL.Image.Image]]


def retrieve_timesteps(scheduler, num_inference_steps: Optional[int] = None, device=None,
                       timesteps: Optional[List[int]] = None, sigmas: Optional[List[float]] = None, **kwargs):
    """cog:95-151: calls ``scheduler.set_timesteps`` (custom ``timesteps`` / ``sigmas`` only if it accepts them)."""
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


class CogVideoXImageToVideoPipeline(DiffusionPipelineBase):
    """Image-to-video generation with CogVideoX + ALG on the native sm_100a kernels (reference class: cog:168-226)."""

    _optional_components = []
    model_cpu_offload_seq = "text_encoder->transformer->vae"
    _callback_tensor_inputs = ["latents", "prompt_embeds", "negative_prompt_embeds"]

    def __init__(self, tokenizer, text_encoder, vae, transformer: CogVideoXTransformer3DModel, scheduler):
        self.register_modules(tokenizer=tokenizer, text_encoder=text_encoder, vae=vae, transformer=transformer,
                              scheduler=scheduler)
        has_vae = getattr(self, "vae", None) is not None
        self.vae_scale_factor_spatial = 2 ** (len(self.vae.config.block_out_channels) - 1) if has_vae else 8
        self.vae_scale_factor_temporal = self.vae.config.temporal_compression_ratio if has_vae else 4
        self.vae_scaling_factor_image = self.vae.config.scaling_factor if has_vae else 0.7
        self.video_processor = VideoProcessor(vae_scale_factor=self.vae_scale_factor_spatial)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, transformer=None, vae=None, torch_dtype=torch.bfloat16,
                        cache_dir=None, synthetic: Optional[bool] = None, allow_synthetic_aux: bool = False,
                        seed: int = 0, device="cuda", native_vae_encoder: bool = True, tokenizer=None, text_encoder=None,
                        **config_overrides):
        """run.py:65-70.  Offline there are no checkpoints: ``synthetic=True`` (or ``ALG_SYNTHETIC=1``) builds the true
        CogVideoX-5b-I2V architecture with seeded random weights directly on ``device``.

        VAE: ``encode`` -- called every step by pixel-space ALG (cog:645) -- runs on the native encoder
        (``alg_b200.vae_cogvideox``): its weights come from the snapshot's ``vae/`` folder (or are synthetic); ``vae=``,
        when given, serves ``decode`` (once per video, not built) -- or everything if ``native_vae_encoder=False``."""
        import os

        if synthetic is None:
            synthetic = os.environ.get("ALG_SYNTHETIC", "0") == "1" or str(pretrained_model_name_or_path).startswith("synthetic")
        scheduler = None
        if not synthetic:
            from alg_b200 import checkpoint

            snap = checkpoint.resolve_snapshot(str(pretrained_model_name_or_path), cache_dir)
            if snap is None:
                raise FileNotFoundError(
                    f"no local diffusers snapshot for {pretrained_model_name_or_path!r}: there is no network, so pass a "
                    "directory (or a hub id already present under cache_dir), or synthetic=True / ALG_SYNTHETIC=1")
            if transformer is None:
                transformer, scheduler = checkpoint.build_from_snapshot(CogVideoXTransformer3DModel, CogVideoXDDIMScheduler, snap, device)
            else:
                scheduler = CogVideoXDDIMScheduler.from_config(checkpoint.scheduler_config(snap))
            if native_vae_encoder and os.path.isdir(os.path.join(snap, "vae")):
                # the snapshot's own VAE on the native kernels: encoder (per-step path of pixel-space ALG) AND decoder; a `vae=`
                # object, when given, keeps serving decode
                vae = AutoencoderKLCogVideoX.from_pretrained(snap, device=device, decoder=vae)
            if vae is None and not allow_synthetic_aux:
                raise NotImplementedError(checkpoint.AUX_MESSAGE)
        if transformer is None:
            transformer = CogVideoXTransformer3DModel.from_synthetic(seed=seed, device=device, **config_overrides)
        if vae is None:
            vae = SyntheticVideoVAE(z_dim=transformer.config.in_channels // 2, scaling_factor=0.7, dtype=torch_dtype)
            if native_vae_encoder and synthetic:
                vae = AutoencoderKLCogVideoX.from_synthetic(seed=seed, device=device, decoder=None, with_decoder=True,
                                                            latent_channels=transformer.config.in_channels // 2)
        if text_encoder is None and not synthetic:  # native T5 v1.1 + the snapshot's tokenizer (cog:228-268)
            from alg_b200 import checkpoint, encoders

            tokenizer, text_encoder = checkpoint.load_text_stack(snap, encoders.T5EncoderModel, device, tokenizer)
            if text_encoder is None and not allow_synthetic_aux:
                raise NotImplementedError(checkpoint.AUX_MESSAGE)
        return cls(tokenizer=tokenizer or SyntheticTokenizer(vocab_size=32128),
                   text_encoder=text_encoder or SyntheticTextEncoder(transformer.config.text_embed_dim, torch_dtype), vae=vae,
                   transformer=transformer, scheduler=scheduler or CogVideoXDDIMScheduler())

    # ------------------------------------------------------------------------------------------------
    # once-per-video conditioning (cog:228-350)
    def _get_t5_prompt_embeds(self, prompt=None, num_videos_per_prompt: int = 1, max_sequence_length: int = 226,
                              device=None, dtype=None):
        device = device or self._execution_device
        dtype = dtype or self.text_encoder.dtype
        prompt = [prompt] if isinstance(prompt, str) else prompt
        batch_size = len(prompt)
        # cog:242-259: tokenizer -> text_encoder(ids)[0] through the transformers call surface (no attention mask: the padded
        # positions take part in T5's self-attention, as in the reference); real HF objects, alg_b200.encoders.T5EncoderModel
        # and the synthetic stand-ins are interchangeable
        text_inputs = self.tokenizer(prompt, padding="max_length", max_length=max_sequence_length, truncation=True,
                                     add_special_tokens=True, return_tensors="pt")
        prompt_embeds = self.text_encoder(text_inputs.input_ids.to(device))[0].to(dtype=dtype, device=device)
        _, seq_len, _ = prompt_embeds.shape
        prompt_embeds = prompt_embeds.repeat(1, num_videos_per_prompt, 1)
        return prompt_embeds.view(batch_size * num_videos_per_prompt, seq_len, -1)

    def encode_prompt(self, prompt, negative_prompt=None, do_classifier_free_guidance: bool = True,
                      num_videos_per_prompt: int = 1, prompt_embeds=None, negative_prompt_embeds=None,
                      max_sequence_length: int = 226, device=None, dtype=None):
        device = device or self._execution_device
        prompt = [prompt] if isinstance(prompt, str) else prompt
        batch_size = len(prompt) if prompt is not None else prompt_embeds.shape[0]
        if prompt_embeds is None:
            prompt_embeds = self._get_t5_prompt_embeds(prompt, num_videos_per_prompt, max_sequence_length, device, dtype)
        if do_classifier_free_guidance and negative_prompt_embeds is None:
            negative_prompt = negative_prompt or ""
            negative_prompt = batch_size * [negative_prompt] if isinstance(negative_prompt, str) else negative_prompt
            if prompt is not None and type(prompt) is not type(negative_prompt):
                raise TypeError(f"`negative_prompt` should be the same type to `prompt`, but got {type(negative_prompt)} !="
                                f" {type(prompt)}.")
            elif batch_size != len(negative_prompt):
                raise ValueError(f"`negative_prompt`: {negative_prompt} has batch size {len(negative_prompt)}, but `prompt`:"
                                 f" {prompt} has batch size {batch_size}. Please make sure that passed `negative_prompt` matches"
                                 " the batch size of `prompt`.")
            negative_prompt_embeds = self._get_t5_prompt_embeds(negative_prompt, num_videos_per_prompt, max_sequence_length,
                                                                device, dtype)
        return prompt_embeds, negative_prompt_embeds

    # ------------------------------------------------------------------------------------------------
    def prepare_latents(self, image: torch.Tensor, batch_size: int = 1, num_channels_latents: int = 16,
                        num_frames: int = 13, height: int = 60, width: int = 90, dtype=None, device=None, generator=None,
                        latents=None):
        """Initial noise + [VAE latent of the image | zero frames] (cog:352-425); layout [B, F, C, H, W]."""
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an effective batch"
                             f" size of {batch_size}. Make sure the batch size matches the length of the generators.")
        num_frames = (num_frames - 1) // self.vae_scale_factor_temporal + 1
        h_lat, w_lat = height // self.vae_scale_factor_spatial, width // self.vae_scale_factor_spatial
        shape = (batch_size, num_frames, num_channels_latents, h_lat, w_lat)
        image = image.unsqueeze(2)  # [B, C, F, H, W]
        if isinstance(generator, list):
            image_latents = [retrieve_latents(self.vae.encode(image[i].unsqueeze(0)), generator[i]) for i in range(batch_size)]
        else:
            image_latents = [retrieve_latents(self.vae.encode(img.unsqueeze(0)), generator) for img in image]
        image_latents = torch.cat(image_latents, dim=0).to(dtype).permute(0, 2, 1, 3, 4)  # [B, F, C, H, W]
        if not self.vae.config.invert_scale_latents:
            image_latents = self.vae_scaling_factor_image * image_latents
        else:
            image_latents = 1 / self.vae_scaling_factor_image * image_latents
        latent_padding = torch.zeros((batch_size, num_frames - 1, num_channels_latents, h_lat, w_lat), device=device, dtype=dtype)
        image_latents = torch.cat([image_latents, latent_padding], dim=1)
        if latents is None:
            latents = randn_tensor(shape, generator=generator, device=device, dtype=dtype)
        else:
            latents = latents.to(device)
        latents = latents * self.scheduler.init_noise_sigma
        return latents, image_latents

    def decode_latents(self, latents: torch.Tensor) -> torch.Tensor:
        latents = latents.permute(0, 2, 1, 3, 4)  # [B, C, F, H, W]
        latents = 1 / self.vae_scaling_factor_image * latents
        return self.vae.decode(latents).sample

    def get_timesteps(self, num_inference_steps, timesteps, strength, device):
        init_timestep = min(int(num_inference_steps * strength), num_inference_steps)
        t_start = max(num_inference_steps - init_timestep, 0)
        return timesteps[t_start * self.scheduler.order:], num_inference_steps - t_start

    def prepare_extra_step_kwargs(self, generator, eta):
        params = set(inspect.signature(self.scheduler.step).parameters.keys())
        extra = {}
        if "eta" in params:
            extra["eta"] = eta
        if "generator" in params:
            extra["generator"] = generator
        return extra

    def check_inputs(self, image, prompt, height, width, negative_prompt, callback_on_step_end_tensor_inputs, latents=None,
                     prompt_embeds=None, negative_prompt_embeds=None):
        """Same ValueErrors, in the same order, as cog:463-524."""
        if not isinstance(image, torch.Tensor) and not isinstance(image, PIL.Image.Image) and not isinstance(image, list):
            raise ValueError("`image` has to be of type `torch.Tensor` or `PIL.Image.Image` or `List[PIL.Image.Image]` but is"
                             f" {type(image)}")
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if callback_on_step_end_tensor_inputs is not None and not all(
                k in self._callback_tensor_inputs for k in callback_on_step_end_tensor_inputs):
            raise ValueError(f"`callback_on_step_end_tensor_inputs` has to be in {self._callback_tensor_inputs}, but found "
                             f"{[k for k in callback_on_step_end_tensor_inputs if k not in self._callback_tensor_inputs]}")
        if prompt is not None and prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `prompt`: {prompt} and `prompt_embeds`: {prompt_embeds}. Please make sure to"
                             " only forward one of the two.")
        elif prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`. Cannot leave both `prompt` and `prompt_embeds` undefined.")
        elif prompt is not None and (not isinstance(prompt, str) and not isinstance(prompt, list)):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")
        if prompt is not None and negative_prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `prompt`: {prompt} and `negative_prompt_embeds`:"
                             f" {negative_prompt_embeds}. Please make sure to only forward one of the two.")
        if negative_prompt is not None and negative_prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `negative_prompt`: {negative_prompt} and `negative_prompt_embeds`:"
                             f" {negative_prompt_embeds}. Please make sure to only forward one of the two.")
        if prompt_embeds is not None and negative_prompt_embeds is not None:
            if prompt_embeds.shape != negative_prompt_embeds.shape:
                raise ValueError("`prompt_embeds` and `negative_prompt_embeds` must have the same shape when passed directly, but"
                                 f" got: `prompt_embeds` {prompt_embeds.shape} != `negative_prompt_embeds`"
                                 f" {negative_prompt_embeds.shape}.")

    def fuse_qkv_projections(self) -> None:
        """cog:527-530.  The native engine always evaluates q|k|v as ONE tensor-core GEMM per block (the weights are
        packed at load time), so there is no slower unfused state to leave; the flag keeps the reference's bookkeeping."""
        self.fusing_transformer = True
        fuse = getattr(self.transformer, "fuse_qkv_projections", None)
        if fuse is not None:
            fuse()

    def unfuse_qkv_projections(self) -> None:
        """cog:533-539: warn when fusion was never requested, else clear the flag."""
        if not getattr(self, "fusing_transformer", False):
            import logging

            logging.getLogger(__name__).warning(
                "The Transformer was not initially fused for QKV projections. Doing nothing.")
            return
        unfuse = getattr(self.transformer, "unfuse_qkv_projections", None)
        if unfuse is not None:
            unfuse()
        self.fusing_transformer = False

    def _prepare_rotary_positional_embeddings(self, height: int, width: int, num_frames: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
        """cog:542-584, CogVideoX 1.0 branch (patch_size_t is None)."""
        cfg = self.transformer.config
        grid_height = height // (self.vae_scale_factor_spatial * cfg.patch_size)
        grid_width = width // (self.vae_scale_factor_spatial * cfg.patch_size)
        base_w, base_h = cfg.sample_width // cfg.patch_size, cfg.sample_height // cfg.patch_size
        crops = get_resize_crop_region_for_grid((grid_height, grid_width), base_w, base_h)
        return get_3d_rotary_pos_embed(embed_dim=cfg.attention_head_dim, crops_coords=crops,
                                       grid_size=(grid_height, grid_width), temporal_size=num_frames, device=device)

    def prepare_lp(self, lp_filter_type, lp_blur_sigma, lp_blur_kernel_size, lp_resize_factor, generator, num_frames,
                   use_low_pass_guidance, lp_filter_in_latent, orig_image_latents, orig_image_tensor):
        """Low-passed copy of the image condition (cog:586-703).

        In-latent: the reference permutes to [B, C, F, H, W], filters and permutes back (cog:684-692); both filters act
        per H x W plane, so filtering the [B, F, C, H, W] tensor in place of the two permuted copies is the same
        arithmetic on the same planes -- one CUDA launch.  Pixel space: filter the RGB frame, VAE-encode it and
        ``sample(generator)`` EVERY call (cog:645; the RNG stream position depends on it), then zero-pad the frames.
        """
        if not use_low_pass_guidance:
            return None
        if not lp_filter_in_latent:
            image_lp = lp_utils.apply_low_pass_filter(orig_image_tensor, filter_type=lp_filter_type, blur_sigma=lp_blur_sigma,
                                                      blur_kernel_size=lp_blur_kernel_size, resize_factor=lp_resize_factor)
            encoded_lp = self.vae.encode(image_lp.unsqueeze(2)).latent_dist.sample(generator=generator)
            if not self.vae.config.invert_scale_latents:
                encoded_lp = self.vae_scaling_factor_image * encoded_lp
            else:
                encoded_lp = 1 / self.vae_scaling_factor_image * encoded_lp
            encoded_lp = encoded_lp.permute(0, 2, 1, 3, 4)
            padded_frames = (num_frames - 1) // self.vae_scale_factor_temporal + 1
            current = encoded_lp.shape[1]
            if padded_frames > current:
                b, _, c, h, w = encoded_lp.shape
                pad = torch.zeros((b, padded_frames - current, c, h, w), device=encoded_lp.device, dtype=encoded_lp.dtype)
                lp_image_latents = torch.cat([encoded_lp, pad], dim=1)
            else:
                lp_image_latents = encoded_lp[:, :padded_frames, ...]
        else:
            lp_image_latents = lp_utils.apply_low_pass_filter(orig_image_latents, filter_type=lp_filter_type,
                                                              blur_sigma=lp_blur_sigma, blur_kernel_size=lp_blur_kernel_size,
                                                              resize_factor=lp_resize_factor)
        assert self.transformer.config.patch_size_t is None  # cog:693-699: temporal-patch prepend is a no-op (quirk q7)
        return lp_image_latents.to(dtype=orig_image_latents.dtype)

    # ------------------------------------------------------------------------------------------------
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

    # ------------------------------------------------------------------------------------------------
    def denoise_step(self, i, t, latents, image_latents, image_tensor, prompt_embeds, negative_prompt_embeds, rope,
                     generator, num_frames, num_inference_steps, alg: Dict[str, Any], guidance_scale: float,
                     use_dynamic_cfg: bool = False, t_back: Optional[int] = None):
        """One iteration of cog:1005-1123 for a single sample: returns (new latents, bf16 noise prediction of all passes).

        ``t`` is the host integer timestep.  Branches and pass order follow the reference exactly."""
        do_cfg = guidance_scale > 1.0
        use_lp = alg["use_low_pass_guidance"]
        if do_cfg and use_lp:
            s = lp_utils.get_lp_strength(
                step_index=i, total_steps=num_inference_steps, lp_strength_schedule_type=alg["lp_strength_schedule_type"],
                schedule_interval_start_time=alg["schedule_interval_start_time"],
                schedule_interval_end_time=alg["schedule_interval_end_time"],
                schedule_linear_start_weight=alg["schedule_linear_start_weight"],
                schedule_linear_end_weight=alg["schedule_linear_end_weight"],
                schedule_linear_end_time=alg["schedule_linear_end_time"], schedule_exp_decay_rate=alg["schedule_exp_decay_rate"])
            two_pass = s == 0
            if alg["lp_strength_schedule_type"] == "exponential" and s < 0.1:  # cog:1031-1032 (quirk q11)
                two_pass = True
            sigma = alg["lp_blur_sigma"] * s
            ksize = alg["lp_blur_kernel_size"] * s if alg["schedule_blur_kernel_size"] else alg["lp_blur_kernel_size"]
            factor = 1.0 - (1.0 - alg["lp_resize_factor"]) * s
            lp = self.prepare_lp(lp_filter_type=alg["lp_filter_type"], lp_blur_sigma=sigma, lp_blur_kernel_size=ksize,
                                 lp_resize_factor=factor, generator=generator, num_frames=num_frames,
                                 use_low_pass_guidance=True, lp_filter_in_latent=alg["lp_filter_in_latent"],
                                 orig_image_latents=image_latents, orig_image_tensor=image_tensor)
            if two_pass:  # cog:1067-1068: both passes see lp_image_latents (the unfiltered tensor when s == 0)
                conds, texts = [lp[0], lp[0]], [negative_prompt_embeds[0], prompt_embeds[0]]
            else:
                conds = [image_latents[0], lp[0], lp[0]]
                texts = [negative_prompt_embeds[0], negative_prompt_embeds[0], prompt_embeds[0]]
        elif do_cfg:
            conds, texts = [image_latents[0], image_latents[0]], [negative_prompt_embeds[0], prompt_embeds[0]]
        else:
            if use_lp:  # the reference leaves `two_pass` undefined here (NameError at cog:1084, quirk q3)
                raise ValueError("use_low_pass_guidance=True needs guidance_scale > 1 (the reference has no unguided ALG branch)")
            conds, texts = [image_latents[0]], [prompt_embeds[0]]
        noise_pred = self.transformer.forward_passes([latents[0]] * len(conds), conds, texts, int(t), rope)
        w = guidance_scale
        if do_cfg and not use_lp and use_dynamic_cfg:  # cog:1105-1108
            w = 1 + guidance_scale * ((1 - math.cos(math.pi * ((num_inference_steps - int(t)) / num_inference_steps) ** 5.0)) / 2)
            self._guidance_scale = w
        if isinstance(self.scheduler, CogVideoXDPMScheduler):  # cog:1113-1122: carries old_pred_original_sample
            latents, self._old_pred_original_sample = self.scheduler.step_cfg(
                noise_pred, w if do_cfg else 1.0, getattr(self, "_old_pred_original_sample", None), int(t), t_back, latents,
                generator)
        else:
            latents = self.scheduler.step_cfg(noise_pred, w if do_cfg else 1.0, int(t), latents)
        return latents, noise_pred

    @torch.no_grad()
    def __call__(
        self,
        image: PipelineImageInput,
        prompt: Optional[Union[str, List[str]]] = None,
        negative_prompt: Optional[Union[str, List[str]]] = None,
        height: Optional[int] = None,
        width: Optional[int] = None,
        num_frames: int = 49,
        num_inference_steps: int = 50,
        timesteps: Optional[List[int]] = None,
        guidance_scale: float = 6.0,
        use_dynamic_cfg: bool = False,
        num_videos_per_prompt: int = 1,
        eta: float = 0.0,
        generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
        latents: Optional[torch.FloatTensor] = None,
        prompt_embeds: Optional[torch.FloatTensor] = None,
        negative_prompt_embeds: Optional[torch.FloatTensor] = None,
        output_type: str = "pil",
        return_dict: bool = True,
        attention_kwargs: Optional[Dict[str, Any]] = None,
        callback_on_step_end: Optional[
            Union[Callable[[int, int, Dict], None], PipelineCallback, MultiPipelineCallbacks]
        ] = None,
        callback_on_step_end_tensor_inputs: List[str] = ["latents"],
        max_sequence_length: int = 226,
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
    ) -> Union[CogVideoXPipelineOutput, Tuple]:
        """Generate a video (cog:727-1158).  Arguments, defaults and return type are those of the reference."""
        if isinstance(callback_on_step_end, (PipelineCallback, MultiPipelineCallbacks)):
            callback_on_step_end_tensor_inputs = callback_on_step_end.tensor_inputs
        cfg = self.transformer.config
        height = height or cfg.sample_height * self.vae_scale_factor_spatial
        width = width or cfg.sample_width * self.vae_scale_factor_spatial
        num_frames = num_frames or cfg.sample_frames
        num_videos_per_prompt = 1  # cog:903 (quirk q13)

        self.check_inputs(image=image, prompt=prompt, height=height, width=width, negative_prompt=negative_prompt,
                          callback_on_step_end_tensor_inputs=callback_on_step_end_tensor_inputs, latents=latents,
                          prompt_embeds=prompt_embeds, negative_prompt_embeds=negative_prompt_embeds)
        self._guidance_scale = guidance_scale
        self._current_timestep = None
        self._attention_kwargs = attention_kwargs
        self._interrupt = False

        if prompt is not None and isinstance(prompt, str):
            batch_size = 1
        elif prompt is not None and isinstance(prompt, list):
            batch_size = len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]
        if batch_size != 1:
            raise NotImplementedError("the native loop runs one sample per GPU (independent samples shard across GPUs)")
        device = self._execution_device
        do_classifier_free_guidance = guidance_scale > 1.0

        prompt_embeds, negative_prompt_embeds = self.encode_prompt(
            prompt=prompt, negative_prompt=negative_prompt, do_classifier_free_guidance=do_classifier_free_guidance,
            num_videos_per_prompt=num_videos_per_prompt, prompt_embeds=prompt_embeds,
            negative_prompt_embeds=negative_prompt_embeds, max_sequence_length=max_sequence_length, device=device)
        prompt_embeds = prompt_embeds.to(device).contiguous()
        if negative_prompt_embeds is not None:
            negative_prompt_embeds = negative_prompt_embeds.to(device).contiguous()

        timesteps, num_inference_steps = retrieve_timesteps(self.scheduler, num_inference_steps, device, timesteps)
        self._num_timesteps = len(timesteps)
        timesteps_host = timesteps.tolist()

        image_tensor = self.video_processor.preprocess(image, height=height, width=width).to(device, dtype=prompt_embeds.dtype)
        latent_channels = cfg.in_channels // 2
        latents, image_latents = self.prepare_latents(image_tensor, batch_size * num_videos_per_prompt, latent_channels,
                                                      num_frames, height, width, prompt_embeds.dtype, device, generator, latents)
        extra_step_kwargs = self.prepare_extra_step_kwargs(generator, eta)
        if extra_step_kwargs.get("eta", 0.0) != 0.0 and not isinstance(self.scheduler, CogVideoXDPMScheduler):
            raise NotImplementedError("eta > 0 (stochastic DDIM) is outside the hot path built here")
        self._old_pred_original_sample = None
        image_rotary_emb = (self._prepare_rotary_positional_embeddings(height, width, latents.size(1), device)
                            if cfg.use_rotary_positional_embeddings else None)

        alg = dict(use_low_pass_guidance=use_low_pass_guidance, lp_filter_type=lp_filter_type,
                   lp_filter_in_latent=lp_filter_in_latent, lp_blur_sigma=lp_blur_sigma,
                   lp_blur_kernel_size=lp_blur_kernel_size, lp_resize_factor=lp_resize_factor,
                   lp_strength_schedule_type=lp_strength_schedule_type, schedule_blur_kernel_size=schedule_blur_kernel_size,
                   schedule_interval_start_time=schedule_interval_start_time,
                   schedule_interval_end_time=schedule_interval_end_time,
                   schedule_linear_start_weight=schedule_linear_start_weight,
                   schedule_linear_end_weight=schedule_linear_end_weight, schedule_linear_end_time=schedule_linear_end_time,
                   schedule_exp_decay_rate=schedule_exp_decay_rate)

        num_warmup_steps = max(len(timesteps) - num_inference_steps * self.scheduler.order, 0)
        with self.progress_bar(total=num_inference_steps) as progress_bar:
            for i, t_host in enumerate(timesteps_host):
                if self.interrupt:
                    continue
                t = timesteps[i]
                self._current_timestep = t
                latents, _ = self.denoise_step(i, t_host, latents, image_latents, image_tensor, prompt_embeds,
                                               negative_prompt_embeds, image_rotary_emb, generator, num_frames,
                                               num_inference_steps, alg, guidance_scale, use_dynamic_cfg,
                                               timesteps_host[i - 1] if i > 0 else None)
                latents = latents.to(prompt_embeds.dtype)
                if callback_on_step_end is not None:
                    scope = dict(latents=latents, prompt_embeds=prompt_embeds, negative_prompt_embeds=negative_prompt_embeds)
                    outputs = callback_on_step_end(self, i, t, {k: scope[k] for k in callback_on_step_end_tensor_inputs})
                    latents = outputs.pop("latents", latents)
                    prompt_embeds = outputs.pop("prompt_embeds", prompt_embeds)
                    negative_prompt_embeds = outputs.pop("negative_prompt_embeds", negative_prompt_embeds)
                if i == len(timesteps) - 1 or ((i + 1) > num_warmup_steps and (i + 1) % self.scheduler.order == 0):
                    progress_bar.update()
        self._current_timestep = None

        if not output_type == "latent":
            video = self.decode_latents(latents)
            video = self.video_processor.postprocess_video(video=video, output_type=output_type)
        else:
            video = latents
        self.maybe_free_model_hooks()
        if not return_dict:
            return (video,)
        return CogVideoXPipelineOutput(frames=video)
