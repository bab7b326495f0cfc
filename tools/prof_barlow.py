"""Profiling helper (not a test): a few BarlowLoss fwd+bwd steps at BASELINE configs[2] (2048 x 8192) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
import torch, ssv_b200
n, d = 2048, 8192
g = torch.Generator().manual_seed(7)
zi = (torch.randn(n, d, generator=g) * 1.2 + 0.3).cuda().requires_grad_(True)
zj = (0.7 * zi.detach().cpu() + 0.3 * torch.randn(n, d, generator=g)).cuda().requires_grad_(True)
fn = ssv_b200.BarlowLoss(False, 0.005)
for _ in range(3):
    zi.grad = None; zj.grad = None
    fn(zi, zj).backward()
torch.cuda.synchronize()
