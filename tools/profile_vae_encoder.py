"""One VAE encode of a 1024x1024 uint8 image (host -> latent posterior), for an ncu launch list; prints the
CUDA-event time per repetition."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffusionkit_b200 as dk  # noqa: E402
from diffusionkit_b200 import ops  # noqa: E402
from diffusionkit_b200.config import VAEEncoderConfig  # noqa: E402
from diffusionkit_b200.weights import init_params, vae_encoder_param_specs  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
size = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
ep = init_params(vae_encoder_param_specs(VAEEncoderConfig()), seed=1, dtype=torch.bfloat16, device=dev)
enc = dk.VAEEncoder(ep)
img = torch.randint(0, 256, (B, size, size, 3), dtype=torch.uint8, device=dev)
for _ in range(reps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ops.launch_count()
    s.record()
    out = enc(img)
    e.record()
    torch.cuda.synchronize()
    print(f"vae encode B={B} {size}x{size}: {s.elapsed_time(e):.2f} ms, {ops.launch_count() - n0} launches, "
          f"hidden {tuple(out.shape)}")
