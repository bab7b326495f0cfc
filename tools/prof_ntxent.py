"""Profiling driver (not a test): a few NT-Xent fwd+bwd steps at BASELINE size for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
import torch
import ssv_b200

n = int(os.environ.get("N", 32768))
steps = int(os.environ.get("STEPS", 3))
g = torch.Generator(device="cuda").manual_seed(0)
zi = torch.randn(n, 128, device="cuda", generator=g).requires_grad_(True)
zj = torch.randn(n, 128, device="cuda", generator=g).requires_grad_(True)
fn = ssv_b200.SimclrLoss(True, 0.5)
for _ in range(steps):
    zi.grad = None; zj.grad = None
    loss = fn(zi, zj)
    loss.backward()
torch.cuda.synchronize()
print("loss", loss.item())
