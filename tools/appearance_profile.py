#!/usr/bin/env python
"""Developer tool (GPU box): where the time of the decoupled-appearance step goes (gof_appearance on torch/cuDNN at the C4 crop).
Prints the top CUDA kernels of forward+backward from torch.profiler and the CUDA-event time of the whole step."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussian-opacity-fields_b200"))
import gof_appearance  # noqa: E402

dev = torch.device("cuda")
if os.environ.get("GOF_APP_CUDNN_BENCH") == "1":
    torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = gof_appearance.AppearanceNetwork(67, 3).to(dev)
emb = (torch.randn(64, device=dev) * 1e-4).requires_grad_(True)
img = torch.rand(3, 1080, 1920, device=dev)
gt = torch.rand(3, 1080, 1920, device=dev)
params = list(net.parameters())


def step():
    rgb = img.detach().requires_grad_(True)
    loss = gof_appearance.l1_loss_appearance(rgb, gt, net, emb)
    return torch.autograd.grad(loss, [rgb, emb] + params)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
print("appearance step (fwd+bwd) ms:", e0.elapsed_time(e1) / 10)
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=90))
