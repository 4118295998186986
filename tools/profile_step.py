"""One full C4 step (4-step FLUX denoise of 4 images + VAE decode) between cudaProfilerStart/Stop, launched kernel by
kernel (CUDA graphs off) — for `ncu --profile-from-start off --metrics gpu__time_duration.sum` launch lists."""
import os
import sys

os.environ["DK_CUDA_GRAPHS"] = "0"
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffusionkit_b200 as dk  # noqa: E402
from diffusionkit_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
pipe = dk.FluxPipeline(w16=True, a16=True, shift=1.0, model_version="argmaxinc/mlx-FLUX.1-schnell", device=dev)
cond, pooled = pipe.synthetic_text_embeddings(n_images=4)
cond, pooled = cond.to(dev), pooled.to(dev)
seeds = [1, 2, 3, 4]


def step():
    latent, _ = pipe.denoise_latents(cond, pooled, num_steps=4, cfg_weight=0.0, latent_size=(128, 128), seed=seeds)
    return pipe._decode(ops.cast_to_16(latent, pipe.activation_dtype), want_u8=True)


step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
