"""torchrun worker of tests/test_gpu_multi.py: rank r denoises sample r (seed 42 + r) with weights broadcast from rank 0 over
NCCL (flat arena), and writes its final latents; rank 0 also writes what ONE process produces for every seed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def sample(pipe, seed, inputs, alg, device):
    g = torch.Generator(device=device).manual_seed(seed)
    lat = torch.randn(inputs["latents"].shape, generator=g, device=device)
    return pipe(image=None, image_embeds=inputs["image_embeds"], prompt_embeds=inputs["prompt_embeds"],
                negative_prompt_embeds=inputs["negative_prompt_embeds"], latents=lat, height=128, width=192, num_frames=9,
                num_inference_steps=4, guidance_scale=5.0, output_type="latent", **alg).frames


def main(out_dir):
    import __graft_entry__ as G
    from alg_b200 import distributed as D, wan
    from alg_b200.schedulers import UniPCMultistepScheduler
    from pipeline_wan_image2video_lowpass import WanImageToVideoPipeline
    rank, world, local = D.env_rank()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    cfg, _, inputs, alg = G.tiny_problem(device)
    full = dict(wan.WAN_I2V_14B, **cfg)
    shapes = {k: (s, torch.float32 if any(f in k for f in wan.FP32_KEYS) else torch.bfloat16) for k, s in wan.parameter_shapes(full).items()}
    sd, arena = D.arena_state_dict(shapes, device)
    if rank == 0:
        for k, v in wan.synthetic_state_dict(full, seed=3, device=device).items():
            sd[k].copy_(v)
    else:
        arena.fill_(0xFF)  # NaN patterns: a rank that missed the broadcast cannot produce finite latents
    D.broadcast_arena(arena)
    model = wan.WanTransformer3DModel(**full).load_state_dict(sd)
    pipe = WanImageToVideoPipeline.from_pretrained("synthetic", transformer=model, synthetic=True)
    pipe.scheduler = UniPCMultistepScheduler.from_config(pipe.scheduler.config, flow_shift=5.0)
    pipe.to(device)
    pipe.set_progress_bar_config(disable=True)
    mine = sample(pipe, D.sample_seed(rank), inputs, alg, device)
    torch.save(mine.cpu(), os.path.join(out_dir, f"rank{rank}.pt"))
    if rank == 0:
        torch.save([sample(pipe, D.sample_seed(r), inputs, alg, device).cpu() for r in range(world)], os.path.join(out_dir, "single.pt"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
