"""Reading diffusers-format checkpoints from a LOCAL directory (the real-checkpoint loader).

The engines name their parameters exactly like the diffusers state dicts, so loading a checkpoint is reading the
``<repo>/transformer/*.safetensors`` shards (+ ``config.json``) and handing the tensors to ``load_state_dict``; the
scheduler comes from ``<repo>/scheduler/scheduler_config.json``.  Used by the three pipelines' ``from_pretrained`` when
``pretrained_model_name_or_path`` is a directory (run.py:46-81 passes a hub id + ``cache_dir``; there is no network here,
so only local snapshots can work).  VAE and the text / image encoders are not built.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional, Tuple

import torch

WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"
INDEX_NAME = WEIGHTS_NAME + ".index.json"


def resolve_snapshot(path: str, cache_dir: Optional[str] = None, require: str = "transformer") -> Optional[str]:
    """A local directory that looks like a diffusers pipeline snapshot (it has the ``require`` component folder), or None.
    Accepts the directory itself or a hub id whose snapshot sits in the HF cache layout under ``cache_dir``
    (models--org--name/snapshots/<rev>/)."""
    if os.path.isdir(path) and os.path.isdir(os.path.join(path, require)):
        return path
    if cache_dir:
        repo = os.path.join(cache_dir, "models--" + path.replace("/", "--"))
        root = os.path.join(repo, "snapshots")
        if os.path.isdir(root):
            revs = sorted(os.listdir(root))
            ref = os.path.join(repo, "refs", "main")  # the revision the hub's `main` pointed at when it was cached
            if os.path.isfile(ref):
                main = open(ref).read().strip()
                if main in revs:
                    revs = [main] + [r for r in revs if r != main]
            for rev in revs:
                cand = os.path.join(root, rev)
                if os.path.isdir(os.path.join(cand, require)):
                    return cand
    return None


def read_config(folder: str, name: str = "config.json") -> dict:
    with open(os.path.join(folder, name)) as f:
        cfg = json.load(f)
    return {k: v for k, v in cfg.items() if not k.startswith("_")}  # drop _class_name / _diffusers_version / ...


def load_safetensors_dir(folder: str, device="cpu") -> Dict[str, torch.Tensor]:
    """All tensors of a (possibly sharded) ``diffusion_pytorch_model`` checkpoint, loaded straight onto ``device``."""
    from safetensors import safe_open

    index = os.path.join(folder, INDEX_NAME)
    if os.path.exists(index):
        with open(index) as f:
            shards = sorted(set(json.load(f)["weight_map"].values()))
    elif os.path.exists(os.path.join(folder, WEIGHTS_NAME)):
        shards = [WEIGHTS_NAME]
    else:
        shards = sorted(f for f in os.listdir(folder) if f.endswith(".safetensors"))
    if not shards:
        raise FileNotFoundError(f"no .safetensors weights under {folder}")
    sd: Dict[str, torch.Tensor] = {}
    for shard in shards:
        with safe_open(os.path.join(folder, shard), framework="pt", device=str(device)) as f:
            for k in f.keys():
                sd[k] = f.get_tensor(k)
    return sd


def load_transformer(snapshot: str, device="cuda") -> Tuple[dict, Dict[str, torch.Tensor]]:
    folder = os.path.join(snapshot, "transformer")
    return read_config(folder), load_safetensors_dir(folder, device)


def load_component(snapshot: str, subfolder: str, device="cuda", key_prefix: Optional[str] = None
                   ) -> Tuple[dict, Dict[str, torch.Tensor]]:
    """(config, state dict) of one component folder of a snapshot (e.g. ``vae``); ``key_prefix`` keeps only the tensors
    whose name starts with it (the native VAE loads ``encoder.*`` and ignores the decoder)."""
    from safetensors import safe_open

    folder = os.path.join(snapshot, subfolder)
    cfg = read_config(folder)
    if key_prefix is None:
        return cfg, load_safetensors_dir(folder, device)
    files = sorted(f for f in os.listdir(folder) if f.endswith(".safetensors"))
    if not files:
        raise FileNotFoundError(f"no .safetensors weights under {folder}")
    sd: Dict[str, torch.Tensor] = {}
    for name in files:
        with safe_open(os.path.join(folder, name), framework="pt", device=str(device)) as f:
            for k in f.keys():
                if k.startswith(key_prefix):
                    sd[k] = f.get_tensor(k)
    return cfg, sd


def scheduler_config(snapshot: str) -> dict:
    folder = os.path.join(snapshot, "scheduler")
    return read_config(folder, "scheduler_config.json") if os.path.isdir(folder) else {}


def save_transformer(snapshot: str, config: dict, state_dict: Dict[str, torch.Tensor], class_name: str,
                     max_shard_bytes: int = 1 << 30) -> None:
    """Write ``<snapshot>/transformer`` in the diffusers layout (sharded safetensors + index).  Used by the tests to build
    synthetic checkpoints; also lets a synthetic model be frozen to disk."""
    from safetensors.torch import save_file

    folder = os.path.join(snapshot, "transformer")
    os.makedirs(folder, exist_ok=True)
    with open(os.path.join(folder, "config.json"), "w") as f:
        json.dump(dict(config, _class_name=class_name, _diffusers_version="0.34.0.dev0"), f, indent=2, default=list)
    shards, cur, size = [], {}, 0
    for k, v in state_dict.items():
        nb = v.numel() * v.element_size()
        if cur and size + nb > max_shard_bytes:
            shards.append(cur)
            cur, size = {}, 0
        cur[k] = v.detach().cpu().contiguous()
        size += nb
    shards.append(cur)
    if len(shards) == 1:
        save_file(shards[0], os.path.join(folder, WEIGHTS_NAME))
        return
    weight_map = {}
    for i, sh in enumerate(shards):
        name = f"diffusion_pytorch_model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
        save_file(sh, os.path.join(folder, name))
        weight_map.update({k: name for k in sh})
    with open(os.path.join(folder, INDEX_NAME), "w") as f:
        json.dump({"metadata": {}, "weight_map": weight_map}, f)


def build_from_snapshot(model_cls, scheduler_cls, snapshot: str, device="cuda"):
    """(transformer, scheduler) of a local diffusers snapshot: real DiT weights into the native engine, the scheduler from
    its ``scheduler_config.json`` (unknown keys are ignored like diffusers' ``from_config`` does)."""
    cfg, sd = load_transformer(snapshot, device)
    if isinstance(cfg.get("patch_size"), list):
        cfg["patch_size"] = tuple(cfg["patch_size"])
    if isinstance(cfg.get("rope_axes_dim"), list):
        cfg["rope_axes_dim"] = tuple(cfg["rope_axes_dim"])
    transformer = model_cls(**cfg).load_state_dict(sd)
    return transformer, scheduler_cls.from_config(scheduler_config(snapshot))


AUX_MESSAGE = ("this snapshot has no loadable VAE / text encoder / image encoder for the native loaders (missing folder, or a "
               "component that is not built natively, SURVEY 8(f)): pass real `vae=` / `text_encoder=` / `image_encoder=` objects (anything "
               "with the transformers / diffusers call surface), or allow_synthetic_aux=True for shape-only stand-ins "
               "(latent-space runs with your own prompt_embeds / image_embeds)")


def load_text_stack(snapshot: str, encoder_cls, device="cuda", tokenizer=None, tokenizer_dir: str = "tokenizer",
                    encoder_dir: str = "text_encoder", **encoder_kwargs):
    """(tokenizer, native text encoder) from ``<snapshot>/tokenizer`` + ``<snapshot>/text_encoder`` (or the ``_2`` folders of
    HunyuanVideo); (tokenizer, None) when the encoder folder is missing.  The tokenizer is transformers' own (host-side text
    processing) unless one is handed in."""
    tdir, edir = os.path.join(snapshot, tokenizer_dir), os.path.join(snapshot, encoder_dir)
    if not os.path.isdir(edir) or (tokenizer is None and not os.path.isdir(tdir)):
        return tokenizer, None
    if tokenizer is None:
        from transformers import AutoTokenizer

        tokenizer = AutoTokenizer.from_pretrained(tdir)
    return tokenizer, encoder_cls.from_pretrained(snapshot, subfolder=encoder_dir, device=device, **encoder_kwargs)


def load_image_stack(snapshot: str, encoder_cls, device="cuda", image_processor=None):
    """(image_processor, native CLIP vision tower) from ``<snapshot>/image_processor`` + ``<snapshot>/image_encoder``."""
    pdir, edir = os.path.join(snapshot, "image_processor"), os.path.join(snapshot, "image_encoder")
    if not os.path.isdir(edir) or (image_processor is None and not os.path.isdir(pdir)):
        return image_processor, None
    if image_processor is None:
        from transformers import CLIPImageProcessor

        image_processor = CLIPImageProcessor.from_pretrained(pdir)
    return image_processor, encoder_cls.from_pretrained(snapshot, device=device)


def save_component(snapshot: str, subfolder: str, config: dict, state_dict: Dict[str, torch.Tensor]) -> None:
    """Write ``<snapshot>/<subfolder>/{config.json, model.safetensors}`` (transformers layout; used by the tests)."""
    from safetensors.torch import save_file

    folder = os.path.join(snapshot, subfolder)
    os.makedirs(folder, exist_ok=True)
    with open(os.path.join(folder, "config.json"), "w") as f:
        json.dump(config, f, indent=2, default=list)
    save_file({k: v.detach().cpu().contiguous() for k, v in state_dict.items()}, os.path.join(folder, "model.safetensors"))
