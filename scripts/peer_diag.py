"""diagnostic: which kinds of access to a CUDA-IPC peer buffer work from this process's kernels"""
import ctypes as C, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
from tinysql_b200 import _lib as L
lib = L.load(); L.check(lib.tq_init(local))
from torch.multiprocessing.reductions import reduce_tensor
dev = torch.device("cuda", local)
buf = torch.zeros(1 << 20, dtype=torch.int64, device=dev)
g = [None] * world
dist.all_gather_object(g, reduce_tensor(buf))
peer = (rank + 1) % world
fn, args = g[peer]
pt = fn(*args)
print(rank, "peer tensor device", pt.device, "ptr", hex(pt.data_ptr()), "can_access", torch.cuda.can_device_access_peer(local, pt.device.index), flush=True)
print(rank, "enable_peer ->", lib.tq_enable_peer_access(pt.device.index), L.last_error(), flush=True)
def step(name, f):
    try:
        f(); lib.tq_device_synchronize(); torch.cuda.synchronize()
        print(rank, name, "OK", flush=True)
    except Exception as e:
        print(rank, name, "FAILED", str(e)[:200], flush=True); raise
step("memset peer via lib", lambda: L.check(lib.tq_memset_device(C.c_void_p(pt.data_ptr()), 1, 64)))
loc = torch.arange(4096, dtype=torch.int64, device=dev)
step("torch copy local->peer", lambda: pt[:4096].copy_(loc))
# push with everything local first, then with the peer as partition 1
k = torch.arange(10000, dtype=torch.int64, device=dev); v = k * 2
def push(dst_ptrs):
    cols = (L.TQColumn * 2)()
    for i, t in enumerate((k, v)):
        cols[i].length, cols[i].data, cols[i].null_bitmap, cols[i].offsets = 10000, t.data_ptr(), None, None
    dest = (C.c_void_p * 4)(*dst_ptrs)
    offs = (C.c_int64 * 2)(0, 0)
    L.check(lib.tq_partition_push_device(2, cols, 0, 10000, 2, dest, offs))
l0, l1, l2, l3 = [torch.zeros(20000, dtype=torch.int64, device=dev) for _ in range(4)]
step("push all-local", lambda: push([l0.data_ptr(), l1.data_ptr(), l2.data_ptr(), l3.data_ptr()]))
dist.barrier()
step("push part1->peer", lambda: push([l0.data_ptr(), l1.data_ptr(), pt.data_ptr(), pt.data_ptr() + 8 * 20000]))
dist.barrier()
print(rank, "received sum", int(buf[:40000].sum()), flush=True)
dist.destroy_process_group()
