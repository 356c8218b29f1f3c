#!/bin/bash
# run.py end to end at the TRUE Wan2.1-I2V-14B-480P architecture with every network on the native kernels (seeded weights: no
# checkpoints offline): SyntheticTokenizer -> UMT5-XXL, CLIP-ViT-H, AutoencoderKLWan encode of the 81-frame condition clip, the ALG
# denoise loop on the 16.4 B-parameter DiT, AutoencoderKLWan decode, post-processing, video file.  STEPS (default 6) shortens the
# schedule to keep the visit short; everything else is configs/wan_alg.yaml.
set -e
STEPS=${1:-6}
OUT=gpurun_out
mkdir -p $OUT
python - <<PY
import yaml
from PIL import Image
c = yaml.safe_load(open("configs/wan_alg.yaml"))
c["generation"]["num_inference_steps"] = $STEPS
yaml.safe_dump(c, open("$OUT/wan_native_demo.yaml", "w"))
im = Image.new("RGB", (832, 480))
px = im.load()
for y in range(480):
    for x in range(832):
        px[x, y] = ((x * 255) // 832, (y * 255) // 480, ((x + y) * 255) // 1312)
im.save("$OUT/wan_native_demo.png")
PY
ALG_SYNTHETIC=1 ALG_NATIVE_ENCODERS=1 ALG_NATIVE_VAE=1 python -c "
import sys, time, types, torch
import run
t0 = time.time()
run.main(types.SimpleNamespace(config='$OUT/wan_native_demo.yaml', image_path='$OUT/wan_native_demo.png',
                               prompt='a red bus turning a corner in the rain', output_path='$OUT/wan_native_demo.mp4', model_cache_dir=None))
torch.cuda.synchronize()
import os
print('run.py wall s', round(time.time() - t0, 1), 'peak GB', round(torch.cuda.max_memory_allocated() / 2**30, 1),
      'mp4 bytes', os.path.getsize('$OUT/wan_native_demo.mp4'))
" 2>&1 | tee $OUT/r02_run_py_native.log | tail -15
