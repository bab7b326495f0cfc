"""Profiling driver (not a test): one EMA step over an 11.3 M-parameter list and one DinoLoss fwd+bwd, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
import torch, ssv_b200 as S
dev = torch.device("cuda", 0)
sizes = [64 * 3 * 9, 64, 64] + [11_200_000 // 60] * 58 + [512 * 1000]
tgt = [torch.randn(n, device=dev) for n in sizes]
src = [torch.randn(n, device=dev) for n in sizes]
up = S.EmaUpdater(tgt, src)
t = torch.randn(1024, 2, 4096, device=dev)
s = torch.randn(1024, 8, 4096, device=dev, requires_grad=True)
c = 0.1 * torch.randn(4096, device=dev)
fn = S.DinoLoss()
for _ in range(3):
    up.step(0.99)
    s.grad = None
    fn(t, s, 0.1, 0.04, c).backward()
torch.cuda.synchronize()
print("ok")
