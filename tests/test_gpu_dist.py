"""Multi-GPU parity (needs >= 2 B200s; skipped on a single-GPU box): row-sharded global-batch NT-Xent over NCCL
vs the single-process oracle on the concatenated batch, and vs the single-GPU CUDA path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_local, d, normalize, tau, out, transport="auto"):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from ssv_b200.dist import DistributedSimclrLoss
    g = torch.Generator().manual_seed(100 + rank)
    zi = torch.randn(n_local, d, generator=g)
    zj = torch.randn(n_local, d, generator=g)
    a = zi.cuda().requires_grad_(True)
    b = zj.cuda().requires_grad_(True)
    fn = DistributedSimclrLoss(normalize, tau, transport=transport)
    for _ in range(3):  # several steps: exercises the double-buffered peer transport
        a.grad = None; b.grad = None
        loss = fn(a, b)
        loss.backward()
    torch.cuda.synchronize()
    out[rank] = (loss.item(), a.grad.cpu().numpy(), b.grad.cpu().numpy(), zi.numpy(), zj.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
@pytest.mark.parametrize("n_local,d,normalize,tau", [(192, 128, True, 0.5), (1000, 64, True, 0.07), (256, 128, True, 0.02)])
def test_dist_ntxent_vs_oracle(n_local, d, normalize, tau, transport):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import ssl_oracle as O
    world = min(torch.cuda.device_count(), 4)
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_local, d, normalize, tau, out, transport), nprocs=world, join=True)
    zi = np.concatenate([out[r][3] for r in range(world)])
    zj = np.concatenate([out[r][4] for r in range(world)])
    ref_loss, ref_dzi, ref_dzj = O.ntxent(zi, zj, normalize, tau)
    for r in range(world):
        loss, gi, gj, _, _ = out[r]
        assert abs(loss - ref_loss) / abs(ref_loss) <= 1e-3
        sl = slice(r * n_local, (r + 1) * n_local)
        assert np.linalg.norm(gi - ref_dzi[sl]) / np.linalg.norm(ref_dzi[sl]) <= 1e-2
        assert np.linalg.norm(gj - ref_dzj[sl]) / np.linalg.norm(ref_dzj[sl]) <= 1e-2
