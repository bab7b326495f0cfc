"""compute-sanitizer driver (not a test): every kernel family once on small shapes.
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
import torch, ssv_b200 as S
from ssv_b200.dist import DistributedBarlowLoss, DistributedMocoLoss, DistributedSwavLoss, DistributedSimclrLoss
F = torch.nn.functional
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
def rn(*s): return torch.randn(*s, device=dev, generator=g)
def run(name, fn, *leaves):
    loss = fn()
    loss.backward()
    torch.cuda.synchronize()
    print(f"{name:28s} loss {loss.item():.5f}")
a, b = rn(200, 96).requires_grad_(True), rn(200, 96).requires_grad_(True)
run("SimclrLoss", lambda: S.SimclrLoss(True, 0.5)(a, b))
run("SimclrLoss raw", lambda: S.SimclrLoss(False, 1.0)(a, b))
run("DistributedSimclrLoss w1", lambda: DistributedSimclrLoss(True, 0.5)(a, b))
q = F.normalize(rn(1000, 96))
run("MocoLoss", lambda: S.MocoLoss(True, 0.07)(a, b, q))
run("DistributedMocoLoss w1", lambda: DistributedMocoLoss(True, 0.07)(a, b, q))
o = rn(200, 96)
run("RelicLoss", lambda: S.RelicLoss(True, 1.0, 0.5)(a, b, o))
run("PirlLoss", lambda: S.PirlLoss(True, 0.07, 0.5)(a, b, F.normalize(o), q))
x, y = rn(136, 264).requires_grad_(True), rn(136, 264).requires_grad_(True)
run("BarlowLoss", lambda: S.BarlowLoss(True, 0.005)(x, y))
run("DistributedBarlow allreduce", lambda: DistributedBarlowLoss(True, 0.005)(x, y))
run("DistributedBarlow colshard", lambda: DistributedBarlowLoss(True, 0.005, mode="colshard")(x, y))
z1, z2 = F.normalize(rn(70, 64)).requires_grad_(True), F.normalize(rn(70, 64)).requires_grad_(True)
for k in (300, 302, 3000):       # TMA-ring fast path (k % 4 == 0) and the general path
    pc = F.normalize(rn(k, 64)).requires_grad_(True)
    bank = F.normalize(rn(50, 64))
    run(f"SwavLoss k={k}", lambda: S.SwavLoss()(z1, z2, pc, bank))
    run(f"DistributedSwavLoss k={k}", lambda: DistributedSwavLoss()(z1, z2, pc, bank))
    codes = S.SwavLoss(0.1, 0.05, 3).compute_codes_sinkhorn((z1.detach() @ pc.detach().t()).contiguous())
    torch.cuda.synchronize()
# round 2: wide rows (KB = 4 kernels), split-K SwAV gradients, the split DinoLoss path, fused-normalise losses, SeLA
aw, bw = rn(150, 200).requires_grad_(True), rn(150, 200).requires_grad_(True)
run("SimclrLoss d=200", lambda: S.SimclrLoss(True, 0.5)(aw, bw))
run("MocoLoss d=200", lambda: S.MocoLoss(True, 0.07)(aw, bw, F.normalize(rn(700, 200))))
z1b, z2b = F.normalize(rn(200, 96)).requires_grad_(True), F.normalize(rn(200, 96)).requires_grad_(True)
pcb = F.normalize(rn(1000, 96)).requires_grad_(True)
run("SwavLoss split-K", lambda: S.SwavLoss()(z1b, z2b, pcb, F.normalize(rn(1000, 96))))
tb, sb = rn(130, 2, 1000), rn(130, 5, 1000).requires_grad_(True)
run("DinoLoss split path", lambda: S.DinoLoss()(tb, sb, 0.1, 0.04, 0.1 * rn(1000)))
run("NormalizedMSELoss", lambda: S.NormalizedMSELoss()(a, b.detach()))
S.SelaLabeler(30, 70, 25.0).step(rn(70, 30), 5)
run("MSELoss", lambda: S.MSELoss()(a, b.detach()))
run("SimSiamLoss", lambda: S.SimSiamLoss()(a, b))
t, s = rn(9, 2, 1000), rn(9, 5, 1000).requires_grad_(True)
run("DinoLoss", lambda: S.DinoLoss()(t, s, 0.1, 0.04, 0.1 * rn(1000)))
c = S.update_teacher_center(None, t.reshape(-1, 1000), 0.9); c = S.update_teacher_center(c, t.reshape(-1, 1000), 0.9)
tg, sr = [rn(n) for n in (7, 8193, 100000)], [rn(n) for n in (7, 8193, 100000)]
S.EmaUpdater(tg, sr).step(0.99)
mb = S.MemoryBank(100, 96); mb.add_batch(a.detach()); mb.add_batch(b.detach())
fb = S.FeatureBank(33, 64); fb.add_vectors(z1.detach())
pb = S.PirlMemoryBank(50, 96); pb.initialize_vectors(torch.arange(20), a.detach()[:20]); pb.update_vectors(torch.arange(20), b.detach()[:20])
pb.get_negatives(torch.tensor([1, 2]))
w = S.Prototypes(64, 30).to(dev); w(dev).sum().backward()
torch.cuda.synchronize()
print("sanitize_small: done")
