"""Ad-hoc GPU debugging helper (not a test): prints per-case errors instead of stopping at the first."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
import numpy as np, torch
import ssv_b200
from oracle import ssl_oracle as O

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))

def run(name, fn):
    try:
        fn()
    except Exception as e:  # noqa
        print(f"[{name}] EXC {type(e).__name__}: {str(e)[:300]}")
        try:
            torch.cuda.synchronize()
        except Exception as e2:
            print("   sync after failure:", str(e2)[:200]); sys.exit(1)

def ntx(n, d, norm, tau):
    g = torch.Generator().manual_seed(0)
    zi, zj = torch.randn(n, d, generator=g), torch.randn(n, d, generator=g)
    a, b = zi.cuda().requires_grad_(True), zj.cuda().requires_grad_(True)
    loss = ssv_b200.SimclrLoss(norm, tau)(a, b)
    torch.cuda.synchronize()
    ref = O.ntxent(zi.numpy(), zj.numpy(), norm, tau)
    print(f"[ntx n={n} d={d} norm={norm} tau={tau}] loss {loss.item():.6f} ref {ref[0]:.6f} rel {abs(loss.item()-ref[0])/abs(ref[0]):.2e}", flush=True)
    loss.backward(); torch.cuda.synchronize()
    print(f"     grads rel {rel(a.grad.cpu().numpy(), ref[1]):.2e} {rel(b.grad.cpu().numpy(), ref[2]):.2e}", flush=True)

if __name__ == "__main__":
    for cfg in [(256, 128, True, 0.5), (64, 64, True, 0.5), (100, 64, True, 0.5), (384, 96, False, 1.0), (2048, 128, True, 0.5), (640, 128, True, 0.02)]:
        run(f"ntx{cfg}", lambda cfg=cfg: ntx(*cfg))
