"""diagnostic: which kinds of access to a CUDA-IPC peer buffer work from this process's kernels"""
import ctypes as C, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
from tinysql_b200 import _lib as L
lib = L.load(); L.check(lib.tq_init(local))
dev = torch.device("cuda", local)
own = C.c_void_p(); L.check(lib.tq_device_alloc(8 << 20, C.byref(own))); L.check(lib.tq_memset_device(own, 0, 8 << 20))
hb = (C.c_ubyte * 64)(); L.check(lib.tq_ipc_get_handle(own, hb))
g = [None] * world
dist.all_gather_object(g, bytes(hb))
peer = (rank + 1) % world
pp = C.c_void_p(); L.check(lib.tq_ipc_open_handle((C.c_ubyte * 64).from_buffer_copy(g[peer]), C.byref(pp)))
class PT:
    def data_ptr(self): return pp.value
pt = PT()
print(rank, "peer ptr", hex(pp.value), flush=True)
def step(name, f):
    try:
        f(); lib.tq_device_synchronize(); torch.cuda.synchronize()
        print(rank, name, "OK", flush=True)
    except Exception as e:
        print(rank, name, "FAILED", str(e)[:200], flush=True); raise
step("memset peer via lib", lambda: L.check(lib.tq_memset_device(C.c_void_p(pt.data_ptr()), 1, 64)))
loc = torch.arange(4096, dtype=torch.int64, device=dev)
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
hostbuf = np.zeros(40000, np.int64); L.check(lib.tq_memcpy_d2h(hostbuf.ctypes.data, own, 320000)); print(rank, "received sum", int(hostbuf.sum()), flush=True)
dist.destroy_process_group()
