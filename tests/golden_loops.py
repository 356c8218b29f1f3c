"""Loader for tests/golden/loop_*.npz (written by oracle/gen_golden_loops.py from the UNMODIFIED reference pipelines)."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(model):
    return sorted(os.path.basename(p)[len("loop_") + len(model) + 1:-4] for p in glob.glob(os.path.join(GOLDEN, f"loop_{model}_*.npz")))


def load(model, name, device="cpu"):
    d = np.load(os.path.join(GOLDEN, f"loop_{model}_{name}.npz"))
    meta = json.loads(bytes(d["meta"]).decode())
    out = {}
    for k in d.files:
        if k == "meta":
            continue
        a = d[k]
        if k.startswith(("text_", "pooled_")) and a.dtype == np.uint8:
            out[k] = bytes(a).decode()
            continue
        t = torch.from_numpy(a.copy())
        if meta["dtypes"].get(k) == "bf16":
            t = t.view(torch.bfloat16)
        out[k] = t.to(device)
    return meta, out


def split_names(s):
    """'n0n1p0p1' / 'nnp' -> ['n0', 'n1', 'p0', 'p1'] / ['n', 'n', 'p']."""
    toks, i = [], 0
    while i < len(s):
        j = i + 1
        while j < len(s) and s[j].isdigit():
            j += 1
        toks.append(s[i:j])
        i = j
    return toks
