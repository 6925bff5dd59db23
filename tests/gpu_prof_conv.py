"""Minimal driver for ncu captures of the tcgen05 kernels at BASELINE config-2 sizes (enc4 fwd / wgrad, B=256)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from srl_zoo_b200 import ops

dev = "cuda"
g = torch.Generator().manual_seed(3)
w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
fpk, dpk = ops.pack_conv_w(w.to(dev), False)
fbf = ops.pack_conv_w_bf16(fpk)
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x = torch.randn(Bn, 56, 56, 64, device=dev)
dy = torch.randn(Bn, 56, 56, 64, device=dev)
out = torch.empty(Bn, 56, 56, 64, device=dev)
for _ in range(3):
    ops.conv64_tc(x, fbf, out, (56, 56), (56, 56), 3, 1, 1, False, want_stats=True)
    ops.wgrad64(x, dy, (56, 56), (56, 56), 3, 1, 1, tensor_cores=True)
torch.cuda.synchronize()
print("done")
