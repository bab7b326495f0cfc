"""Profiling driver (not a test): Sinkhorn at cfg4 for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
import torch, ssv_b200
g = torch.Generator(device="cuda").manual_seed(0)
z = torch.nn.functional.normalize(torch.randn(4096, 128, device="cuda", generator=g), dim=-1)
c = torch.nn.functional.normalize(torch.randn(3000, 128, device="cuda", generator=g), dim=-1)
s = (z @ c.t()).contiguous()
fn = ssv_b200.SwavLoss(0.1, 0.05, 3)
for _ in range(3):
    q = fn.compute_codes_sinkhorn(s)
torch.cuda.synchronize()
print(q.sum().item())
