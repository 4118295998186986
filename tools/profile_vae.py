"""One VAE decode at latent 128x128 (1024x1024 image), for an ncu launch list; prints the CUDA-event time."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffusionkit_b200 as dk  # noqa: E402
from diffusionkit_b200 import ops  # noqa: E402
from diffusionkit_b200.config import VAEDecoderConfig  # noqa: E402
from diffusionkit_b200.weights import init_params, vae_decoder_param_specs  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
vp = init_params(vae_decoder_param_specs(VAEDecoderConfig()), seed=1, dtype=torch.bfloat16, device=dev)
dec = dk.VAEDecoder(vp)
z = torch.randn((B, 128, 128, 16), device=dev).to(torch.bfloat16)
for r in range(reps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    last = r == reps - 1 and os.environ.get("DK_PROFILE_RANGE") == "1"   # ncu --profile-from-start off: last decode only
    if last:
        torch.cuda.profiler.start()
    s.record()
    out = dec(z)
    e.record()
    torch.cuda.synchronize()
    if last:
        torch.cuda.profiler.stop()
    print(f"vae decode B={B}: {s.elapsed_time(e):.2f} ms, launches so far {ops.launch_count()}")
