"""Driver for ncu captures of one engine step (B given on argv)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import srl_zoo_b200

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(1)
mod = srl_zoo_b200.B200SRLModules(200, 6, True, "custom_cnn", ["autoencoder"]).cuda()
eng = srl_zoo_b200.TrainStep(mod, bs)
obs = torch.randn(bs, 3, 224, 224, device="cuda")
nobs = torch.randn(bs, 3, 224, 224, device="cuda")
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    eng.step(obs, nobs)
torch.cuda.synchronize()
print("done")
