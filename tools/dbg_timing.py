"""Timing-experiment driver (needs a build with -DSSVB_DBG_TIMING): per-phase cycle breakdown of sim_bwd_kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
import torch
dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
os.environ["SSVB_DBG_PTR"] = hex(dbg.data_ptr())
import ssv_b200
n = 32768
g = torch.Generator(device="cuda").manual_seed(0)
zi = torch.randn(n, 128, device="cuda", generator=g).requires_grad_(True)
zj = torch.randn(n, 128, device="cuda", generator=g).requires_grad_(True)
fn = ssv_b200.SimclrLoss(True, 0.5)
for _ in range(3):
    zi.grad = None; zj.grad = None
    fn(zi, zj).backward()
torch.cuda.synchronize()
v = dbg.cpu().tolist()
names = ["loop/other", "wait b_full", "wait s_full", "tmem ld+release", "weights[0:32]", "wait w_empty", "weights[32:64]+st", "st drain+arrive"]
tiles = 512 * 512 / 148 / 2  # tiles handled by one warpgroup pair of one CTA (approx.)
tot = sum(v[:8])
print(f"math warp: total {tot} cycles, ~{tot / tiles:.0f} cycles per own tile ({tiles:.0f} tiles)")
for nm, c in zip(names, v[:8]):
    print(f"  {nm:22s} {c:12d}  {100 * c / max(tot, 1):5.1f}%  {c / tiles:7.0f} cyc/tile")
print(f"mma warp: issuing {v[8]} cycles, idle-polling {v[9]} cycles ({100 * v[9] / max(v[8] + v[9], 1):.1f}% idle)")

f = v[16:24]
ftiles = 512 * 256 / 148 / 2
ftot = sum(f)
print(f"fwd softmax warp: total {ftot} cycles, ~{ftot / ftiles:.0f} cycles per own tile ({ftiles:.0f} tiles)")
for nm, c in zip(["loop/other", "wait s_full", "wait tmem ld", "exp/sum compute"], f[:4]):
    print(f"  {nm:22s} {c:12d}  {100 * c / max(ftot, 1):5.1f}%  {c / ftiles:7.0f} cyc/tile")
