"""Probe: does torch symmetric memory give working peer pointers on this box? (2+ ranks via torchrun)"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t = sm.empty(1024, dtype=torch.float32, device="cuda:%d" % lr)
t.fill_(float(rank))
hdl = sm.rendezvous(t, dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal", len(hdl.signal_pad_ptrs), flush=True)
peer = (rank + 1) % world
pb = hdl.get_buffer(peer, (1024,), torch.float32)
hdl.barrier()
pb[rank * 4:(rank + 1) * 4] = 100.0 + rank          # store into the peer's memory
torch.cuda.synchronize()
hdl.barrier()
torch.cuda.synchronize()
print(rank, "local after peer write:", t[:8].tolist(), flush=True)
dist.destroy_process_group()
