"""Profiling / timing driver (not a test): Sinkhorn at cfg4.  `python tools/prof_sinkhorn.py time` prints the
CUDA-graph replay time (L2 flushed between replays) of compute_codes_sinkhorn and of the full SwAV loss."""
import os, sys, statistics
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
if len(sys.argv) > 1 and sys.argv[1] == "time":
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            fn.compute_codes_sinkhorn(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn.compute_codes_sinkhorn(s)
    for warm in (True, False):
        ts = []
        for _ in range(30):
            if not warm:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr.replay(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"sinkhorn 4096x3000 graph replay, {'L2-warm' if warm else 'L2 flushed'}: median {statistics.median(ts):.1f} us, min {min(ts):.1f} us")
