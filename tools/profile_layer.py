"""ncu target: one latent-tower layer (SelfAttention block) and one encoder cross-attend at the ImageNet-recipe
shapes, launched eagerly a few times.  Run plain first, then under `ncu --set full -k regex:pio_`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import perceiverio_pytorch_b200 as pio  # noqa: E402

torch.manual_seed(0)
B = int(os.environ.get("PIO_PROFILE_BATCH", "64"))
layer = pio.SelfAttention(in_channels=1024, widening_factor=1, num_heads=8).eval().cuda()
x = torch.randn(B, 512, 1024, device="cuda")
enc = pio.PerceiverEncoder(num_input_channels=261, num_self_attends_per_block=1, num_blocks=1, num_latents=512,
                           num_latent_channels=1024).eval().cuda()
inputs = torch.randn(8, 50176, 261, device="cuda")
with torch.inference_mode():
    for _ in range(3):
        y = layer(x)
    for _ in range(2):
        z = enc.cross_attend(enc.latents(inputs), inputs)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()), float(z.abs().mean()))
