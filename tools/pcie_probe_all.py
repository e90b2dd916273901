"""What the box gives for plain host-to-device copies with k of its GPUs loaded at once (k = 1, 2, 4, 8), from ordinary
pinned memory and from write-combined pinned memory.  Launch with torch.distributed.run, one rank per GPU.  This is the
ceiling of bench.py's end-to-end arm (host buffers, copies inside the timed region): no library code is involved."""
import ctypes
import json
import os

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
N = 4 * 1024**3
d = torch.empty(N, dtype=torch.uint8, device="cuda")
bufs = {}
for name, flags in (("pinned", 0), ("write_combined", 4)):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), N, flags)
    bufs[name] = p if rc == 0 else None
    if rc == 0:
        ctypes.memset(p, 1, N)
out = {}
ks = [k for k in (1, 2, 4, 8) if k <= world]
for name, p in bufs.items():
    for k in ks:
        dist.barrier()
        torch.cuda.synchronize()
        gbs = 0.0
        if rank < k and p is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            rt.cudaMemcpyAsync(d.data_ptr(), p, N, 1, None)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                rt.cudaMemcpyAsync(d.data_ptr(), p, N, 1, None)
            e1.record()
            torch.cuda.synchronize()
            gbs = 3 * N / (e0.elapsed_time(e1) * 1e-3) / 1e9
        t = torch.tensor([gbs], device="cuda")
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        if rank == 0:
            per = [round(float(x.item()), 1) for x in allv[:k]]
            out["%s_k%d" % (name, k)] = {"per_gpu_GBps": per, "aggregate_GBps": round(sum(per), 1)}
if rank == 0:
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
