#!/usr/bin/env python
"""Does torch's symmetric memory (CUDA VMM peer mappings over NVLink) work on this box?
Run under torchrun with >= 2 ranks.  Prints one line per rank."""
import os
import time

import torch
import torch.distributed as td


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    td.init_process_group("nccl", device_id=dev)
    import torch.distributed._symmetric_memory as symm_mem

    t = symm_mem.empty(1 << 20, dtype=torch.uint8, device=dev)
    t.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hdl = symm_mem.rendezvous(t, td.group.WORLD)
    dt = time.perf_counter() - t0
    ptrs = list(hdl.buffer_ptrs)
    # every rank writes its rank + 1 into byte `rank` of EVERY peer's buffer through the peer mapping
    for r in range(world):
        peer = hdl.get_buffer(r, (1 << 20,), torch.uint8, 0)
        peer[rank] = rank + 1
    torch.cuda.synchronize()
    td.barrier()
    torch.cuda.synchronize()
    got = t[:world].tolist()
    print(f"rank {rank}: rendezvous {1e3 * dt:.1f} ms, ptrs {[hex(p) for p in ptrs]}, multicast_ptr {hex(hdl.multicast_ptr) if getattr(hdl, 'multicast_ptr', 0) else None}, "
          f"bytes {got}, ok {got == list(range(1, world + 1))}", flush=True)
    td.destroy_process_group()


if __name__ == "__main__":
    main()
