"""2-rank probe: does torch symmetric memory work on this box, and what does the handle expose?"""
import os, sys, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as sm
try:
    t = sm.empty(1 << 20, dtype=torch.uint8, device=torch.device("cuda", local))
    h = sm.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok; attrs:", [a for a in dir(h) if not a.startswith("_")])
    print(rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in h.signal_pad_ptrs], "rank", h.rank, "world", h.world_size)
    t.zero_(); torch.cuda.synchronize(); dist.barrier()
    # write my rank into the peer's buffer through the mapped pointer
    peer = h.get_buffer((rank + 1) % world, (16,), torch.uint8)
    peer.fill_(rank + 1); torch.cuda.synchronize(); dist.barrier()
    print(rank, "my buffer now holds", t[:4].tolist())
except Exception as e:
    print(rank, "symmetric memory FAILED:", repr(e))
dist.destroy_process_group()
