"""Timing-experiment driver (needs a build with -DSSVB_DBG_TIMING, SSVB_LIB=libssv_b200_dbg.so): per-phase cycle
breakdown of the fused single-pass MoCo kernel (sim_bwd_kernel<2, SIM_MOCO>) for one math warp of CTA 0."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
import torch
dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
os.environ["SSVB_DBG_PTR"] = hex(dbg.data_ptr())
import ssv_b200 as S
n, k, d = 256, 65536, 128
bank = S.MemoryBank(k, d)
fill = torch.randn(k, d, device="cuda")
for i in range(4):
    bank.add_batch(fill[i * 16384:(i + 1) * 16384])
a = torch.randn(n, d, device="cuda", requires_grad=True)
b = torch.randn(n, d, device="cuda", requires_grad=True)
fn = S.MocoLoss(True, 0.07)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    flush.zero_()
    dbg.zero_()
    a.grad = None; b.grad = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(a, b, bank.get_vectors()).backward()
    e1.record()
    torch.cuda.synchronize()
v = dbg.cpu().tolist()
names = ["loop/other", "wait b_full", "wait s_full", "tmem ld+release", "weights[0:64]", "wait w_empty", "tmem st issue", "st drain+arrive"]
tot = sum(v[:8])
print(f"eager fwd+bwd {e0.elapsed_time(e1) * 1e3:.1f} us; math warp of CTA 0 (tile loop only): {tot} cycles = {tot / 1.9e3:.2f} us @1.9 GHz")
for nm, c in zip(names, v[:8]):
    print(f"  {nm:22s} {c:10d}  {100 * c / max(tot, 1):5.1f}%")
print(f"S-MMA warp: issuing {v[8]} cycles, idle-polling {v[9]} cycles")
