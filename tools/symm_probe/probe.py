"""2-GPU probe of torch symmetric memory (peer pointers over NVLink) - bring-up for the fused encoder -> all-gather."""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
buf = symm_mem.empty(world * 4, 8, dtype=torch.float32, device=f"cuda:{lr}")
hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "attrs", [a for a in dir(hdl) if not a.startswith("_")][:40], flush=True)
buf.zero_()
hdl.barrier()
for peer in range(world):
    dst = hdl.get_buffer(peer, (world * 4, 8), torch.float32)
    dst[rank * 4:(rank + 1) * 4] = float(rank + 1)
hdl.barrier()
torch.cuda.synchronize()
print(rank, "rows", buf[:, 0].tolist(), flush=True)
dist.destroy_process_group()
