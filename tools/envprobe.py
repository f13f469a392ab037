import os
keys = sorted(k for k in os.environ if any(s in k for s in ("INJECT", "NV_", "NSIGHT", "NCU", "CUDA", "PRELOAD", "NVTX")))
open("gpurun_out/env_under_ncu.txt", "a").write(repr({k: os.environ[k][:80] for k in keys}) + "\n")
import torch
torch.zeros(4, device="cuda").sum().item()
