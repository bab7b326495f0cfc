"""Probe (not a test): does torch symmetric memory (peer pointers over NVLink) work on this box?"""
import os, sys
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dev = torch.device("cuda", lr_)
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty(1024, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs][:2],
      "attrs", [a for a in dir(hdl) if not a.startswith("_")])
t.fill_(float(rank + 1))
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
print(rank, "peer value", peer[0].item())
hdl.barrier()
dist.destroy_process_group()
