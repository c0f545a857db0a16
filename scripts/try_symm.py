"""Probe torch symmetric memory on this box (run under torchrun)."""
import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
import torch.distributed._symmetric_memory as symm_mem
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
try:
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok", type(hdl).__name__, "ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], "signal pads", len(hdl.signal_pad_ptrs),
          "pad size", getattr(hdl, "signal_pad_size", None), "multicast", hex(getattr(hdl, "multicast_ptr", 0) or 0))
    t.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (16,), torch.float32)
    print(rank, "peer value", peer[:2].tolist())
    hdl.barrier()
    print(rank, [a for a in dir(hdl) if not a.startswith("_")])
except Exception as e:
    import traceback; traceback.print_exc()
dist.barrier(); dist.destroy_process_group()
