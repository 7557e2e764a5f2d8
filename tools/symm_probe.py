"""Probe: does torch symmetric memory (peer pointers over NVLink) work on this box?  Run under torchrun with 2+ ranks."""
import os
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.int64, device="cuda")
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "symm ok", type(hdl).__name__, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], "signal_pad_ptrs", len(hdl.signal_pad_ptrs), flush=True)
    t.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (16,), torch.int64)
    print(rank, "peer read", peer[:4].tolist(), flush=True)
    hdl.barrier()
except Exception as e:
    print(rank, "symm FAILED", repr(e), flush=True)
print(rank, "can_access_peer", [torch.cuda.can_device_access_peer(lr, j) for j in range(world) if j != lr], flush=True)
dist.destroy_process_group()
