"""run.py -- same CLI, YAML schema and control flow as the reference driver (run.py:26-146), on the B200-native engine.

    python run.py --config configs/wan_alg.yaml --image_path img.png --prompt "..." [--output_path out.mp4]
                  [--model_cache_dir DIR]

`diffusers` is not a dependency: the pipeline classes come from this repo's drop-in modules (same module and class
names as the reference's imports at run.py:15-19) and the scheduler classes from ``alg_b200.schedulers``.  Offline
there are no checkpoints: with ``ALG_SYNTHETIC=1`` the ``model.path`` of the config only selects the architecture
(seeded random weights; ``ALG_NATIVE_ENCODERS=1`` / ``ALG_NATIVE_VAE=1`` add seeded encoders / VAE at their true architectures
instead of the shape-only stand-ins); with a local diffusers snapshot (a directory, or the hub id under ``--model_cache_dir``)
every component -- DiT, scheduler config, text / image encoders, VAE -- loads into the native engines like run.py:46-81 does;
without either ``from_pretrained`` raises FileNotFoundError.
"""
import argparse
import logging
import sys

import torch
import yaml
from PIL import Image

from alg_b200.pipeline_utils import load_image, write_video
from alg_b200.schedulers import FlowMatchEulerDiscreteScheduler, UniPCMultistepScheduler
from lp_utils import get_hunyuan_video_size
from pipeline_cogvideox_image2video_lowpass import CogVideoXImageToVideoPipeline
from pipeline_hunyuan_video_image2video_lowpass import HunyuanVideoImageToVideoPipeline
from pipeline_wan_image2video_lowpass import WanImageToVideoPipeline

logging.basicConfig(level=logging.INFO, format='%(asctime)s - %(levelname)s - %(message)s', stream=sys.stdout)
logger = logging.getLogger(__name__)


def main(args):
    # 1. Configuration
    with open(args.config, 'r') as f:
        config = yaml.safe_load(f)
    model_path = config['model']['path']
    model_dtype = getattr(torch, config['model']['dtype'])
    device = "cuda" if torch.cuda.is_available() else "cpu"
    logger.info(f"Using device: {device}")
    if device != "cuda":
        raise RuntimeError("the ALG engine is sm_100a CUDA only: there is no CPU fallback")

    # 2. Pipeline preparation (run.py:45-87)
    from alg_b200 import checkpoint
    have_snapshot = checkpoint.resolve_snapshot(str(model_path), args.model_cache_dir) is not None  # else: ALG_SYNTHETIC=1 (seeded weights)
    if "Wan" in model_path:
        extra = {}
        if have_snapshot:  # run.py:46-55: the image encoder and the VAE are loaded in float32 and handed to the pipeline
            from alg_b200.encoders import CLIPVisionModel
            from alg_b200.vae_wan import AutoencoderKLWan
            extra["image_encoder"] = CLIPVisionModel.from_pretrained(model_path, subfolder="image_encoder", torch_dtype=torch.float32,
                                                                     cache_dir=args.model_cache_dir)
            extra["vae"] = AutoencoderKLWan.from_pretrained(model_path, subfolder="vae", torch_dtype=torch.float32,
                                                            cache_dir=args.model_cache_dir)
        pipe = WanImageToVideoPipeline.from_pretrained(model_path, torch_dtype=model_dtype, cache_dir=args.model_cache_dir, **extra)
        # run.py:63 compares the YAML int 480 with the string '480', so flow_shift is always 5.0 (quirk q1): kept as shipped
        pipe.scheduler = UniPCMultistepScheduler.from_config(
            pipe.scheduler.config, flow_shift=3.0 if config['generation']['height'] == '480' else 5.0)
    elif "CogVideoX" in model_path:
        pipe = CogVideoXImageToVideoPipeline.from_pretrained(model_path, torch_dtype=model_dtype, cache_dir=args.model_cache_dir)
    elif "HunyuanVideo" in model_path:
        extra = {}
        if have_snapshot:  # run.py:71-76: the transformer is loaded in bfloat16 on its own, the rest of the pipeline in float16
            from alg_b200.hunyuan import HunyuanVideoTransformer3DModel
            extra["transformer"] = HunyuanVideoTransformer3DModel.from_pretrained(model_path, subfolder="transformer",
                                                                                  torch_dtype=torch.bfloat16, cache_dir=args.model_cache_dir)
        pipe = HunyuanVideoImageToVideoPipeline.from_pretrained(model_path, torch_dtype=torch.float16,
                                                                cache_dir=args.model_cache_dir, **extra)
        pipe.scheduler = FlowMatchEulerDiscreteScheduler.from_config(
            pipe.scheduler.config, flow_shift=config['model']['flow_shift'], invert_sigmas=config['model']['flow_reverse'])
    else:
        raise ValueError(f"unknown model family in model.path: {model_path!r}")
    pipe.to(device)
    logger.info("Pipeline loaded successfully.")

    # 3. Prepare inputs
    input_image = load_image(Image.open(args.image_path))
    generator = torch.Generator(device=device).manual_seed(42)
    pipe_kwargs = {"image": input_image, "prompt": args.prompt, "generator": generator}
    params_from_config = {**config.get('generation', {}), **config.get('alg', {})}
    for key, value in params_from_config.items():
        if value is not None:
            pipe_kwargs[key] = value
    logger.info("Starting video generation...")
    log_subset = {k: v for k, v in pipe_kwargs.items() if k not in ['image', 'generator']}
    logger.info(f"Pipeline arguments: {log_subset}")
    if "HunyuanVideo" in model_path:
        pipe_kwargs["height"], pipe_kwargs["width"] = get_hunyuan_video_size(config['video']['resolution'], input_image)

    # 4. Generate video
    video_output = pipe(**pipe_kwargs)
    video_frames = video_output.frames[0]
    logger.info(f"Video generation complete. Received {len(video_frames)} frames.")

    # 5. Save video (run.py:121-133; torchvision.io.write_video is gone from this image's torchvision: cv2 writer)
    logger.info(f"Saving video to: {args.output_path}")
    write_video(args.output_path, video_frames, fps=config['video']['fps'])
    logger.info("Video saved successfully. Run complete.")


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Arguments")
    parser.add_argument("--config", type=str, default="./configs/hunyuan_video_alg.yaml")
    parser.add_argument("--image_path", type=str, default="./assets/a red double decker bus driving down a street.jpg")
    parser.add_argument("--prompt", type=str, default="a red double decker bus driving down a street")
    parser.add_argument("--output_path", type=str, default="output.mp4")
    parser.add_argument("--model_cache_dir", type=str, default=None)
    args = parser.parse_args()
    main(args)
