"""Multi-GPU hash join: radix-partition both sides on the join key across the ranks, exchange the partitions with ONE
grouped NCCL send/recv (the all-to-all at the shard boundary — NCCL 2.27/2.28 has no ncclAllToAllv), then run the
single-GPU join on what arrived.  One process per GPU; torch.distributed is plumbing only (device buffers + the
collective); partitioning and joining are the CUDA kernels of libtinysql_b200.so.

The moral equivalent of the reference's partial->final hash shuffle (executor/aggregate.go:96-133,352-356): a row goes to
rank  (mix64(key) >> 40) % world  — a pure function of its key, so equal keys meet on one rank and nothing else moves.

The exchange logic is backend-agnostic (gloo on CPU in tests/test_dist_gloo.py with injected partition / join functions).
"""
import ctypes as C
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def mix64_np(k):
    """tqd::mix64 (csrc/common.cuh) in numpy, for the CPU stand-in partitioner of the gloo tests."""
    k = k.astype(np.uint64).copy()
    k ^= k >> np.uint64(33)
    k *= np.uint64(0xFF51AFD7ED558CCD)
    k ^= k >> np.uint64(33)
    k *= np.uint64(0xC4CEB9FE1A85EC53)
    k ^= k >> np.uint64(33)
    return k


def dest_rank_np(keys, world):
    return ((mix64_np(keys) >> np.uint64(40)) % np.uint64(world)).astype(np.int64)


def exchange_counts(send_offsets_list, world, rank, device, group=None):
    """One all-gather for any number of partitioned tables: returns, per table, the rows every source rank sends HERE."""
    mine = torch.tensor([[offs[p + 1] - offs[p] for p in range(world)] for offs in send_offsets_list], dtype=torch.int64, device=device)
    allc = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine, group=group)            # world x tables x world counts (tiny)
    m = torch.stack(allc).cpu().numpy()                 # ONE device->host sync
    return [[int(m[src, t, rank]) for src in range(world)] for t in range(len(send_offsets_list))]


def exchange(cols, send_offsets, world, rank, group=None, recv_counts=None, async_op=False):
    """cols: list of 1-D tensors, all partitioned the same way: rows [send_offsets[p], send_offsets[p+1]) go to rank p.
    Returns (list of received columns, recv_counts[, work handles]).  NCCL: one all_to_all_single per column (NCCL
    send/recv groups inside); other backends (gloo has no all-to-all): one grouped batch of isend/irecv."""
    dev = cols[0].device
    if recv_counts is None:
        recv_counts = exchange_counts([send_offsets], world, rank, dev, group)[0]
    send_counts = [int(send_offsets[p + 1] - send_offsets[p]) for p in range(world)]
    total = int(sum(recv_counts))
    out = [torch.empty(total, dtype=c.dtype, device=dev) for c in cols]
    works = []
    if dist.get_backend(group) == "nccl":
        for ci, c in enumerate(cols):
            src = c[int(send_offsets[0]): int(send_offsets[world])]
            w = dist.all_to_all_single(out[ci], src, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=group, async_op=async_op)
            if async_op:
                works.append(w)
    else:
        recv_off = np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64)
        ops = []
        for ci, c in enumerate(cols):
            for peer in range(world):
                s_lo, s_hi = int(send_offsets[peer]), int(send_offsets[peer + 1])
                r_lo, r_hi = int(recv_off[peer]), int(recv_off[peer + 1])
                if peer == rank:
                    out[ci][r_lo:r_hi].copy_(c[s_lo:s_hi])
                    continue
                if s_hi > s_lo:
                    ops.append(dist.P2POp(dist.isend, c[s_lo:s_hi], peer, group=group))
                if r_hi > r_lo:
                    ops.append(dist.P2POp(dist.irecv, out[ci][r_lo:r_hi], peer, group=group))
        if ops:
            for req in dist.batch_isend_irecv(ops):  # one group around every send and recv
                req.wait()
    if async_op:
        return out, recv_counts, works
    return out, recv_counts


class PeerExchange:
    """All-to-all over NVSwitch PEER MEMORY instead of NCCL send/recv: every rank keeps its partitioned columns in
    persistent "send" buffers whose CUDA IPC handles are shared once; an exchange is then one tiny offsets all-gather
    plus, per source rank, a peer-to-peer copy (copy engines over NVLink) that PULLS the slice addressed to this rank.
    Two barriers per exchange order the producers' scatter kernels against the consumers' pulls."""

    def __init__(self, world, rank, device, n_tables_cols, capacity_rows, dtype=torch.int64):
        from torch.multiprocessing.reductions import reduce_tensor
        self.world, self.rank, self.dev = world, rank, device
        # send[t][c]: partition output of table t, column c (written by tq_partition_device)
        self.send = [[torch.empty(cap, dtype=dtype, device=device) for _ in range(nc)] for nc, cap in zip(n_tables_cols, capacity_rows)]
        handles = [[reduce_tensor(x) for x in cols] for cols in self.send]
        gathered = [None] * world
        dist.all_gather_object(gathered, handles)
        self.peer = []
        for src in range(world):
            if src == rank:
                self.peer.append(self.send)
            else:
                self.peer.append([[fn(*args) for fn, args in cols] for cols in gathered[src]])
        self.streams = [torch.cuda.Stream(device=device) for _ in range(max(1, min(world - 1, 4)))]

    def offsets_matrix(self, offsets_list):
        """all ranks' partition offsets for every table: [src][table][world+1] (one all-gather, one D2H sync)"""
        mine = torch.tensor(offsets_list, dtype=torch.int64, device=self.dev)
        allo = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allo, mine)
        return torch.stack(allo).cpu().numpy()

    def pull(self, table, offs, out=None):
        """copy, from every source rank, the rows it partitioned for THIS rank.  offs = offsets_matrix(...)[:, table, :]"""
        world, rank = self.world, self.rank
        counts = [int(offs[src, rank + 1] - offs[src, rank]) for src in range(world)]
        total = sum(counts)
        ncols = len(self.send[table])
        if out is None:
            out = [torch.empty(total, dtype=self.send[table][c].dtype, device=self.dev) for c in range(ncols)]
        cur = torch.cuda.current_stream()
        pos = 0
        events = []
        for i, src in enumerate([(rank + d) % world for d in range(world)]):  # start with the local slice, then ring order
            lo, hi = int(offs[src, rank]), int(offs[src, rank + 1])
            dst_lo = sum(counts[:src])
            st = self.streams[i % len(self.streams)]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                for c in range(ncols):
                    out[c][dst_lo:dst_lo + (hi - lo)].copy_(self.peer[src][table][c][lo:hi], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(st)
            events.append(ev)
        return out, counts, events


class RawCol:
    """a device column that is just (pointer, rows): quacks like the torch tensors the helpers below take"""

    def __init__(self, ptr, n):
        self.ptr, self.n = ptr, n

    def data_ptr(self):
        return self.ptr

    def numel(self):
        return self.n

    def __getitem__(self, sl):
        assert sl.start in (None, 0) and sl.step in (None, 1)
        return RawCol(self.ptr, min(self.n, sl.stop))


class PushExchange:
    """The fused scatter + exchange: every rank owns persistent RECEIVE buffers (tq_device_alloc) exported once as CUDA IPC
    handles; every peer opens them under its own device (lazy peer access, like NCCL's P2P transport).  Per step each
    rank counts its rows per destination (tq_partition_count_device), one tiny all-gather turns the counts into write
    offsets, and tq_partition_push_device scatters every row straight into the destination rank's receive buffer —
    stores over NVLink peer memory, no separate exchange pass.  Barriers order pushes against consumers."""

    def __init__(self, lib, L, world, rank, device, n_tables_cols, capacity_rows):
        self.lib, self.L, self.world, self.rank, self.dev = lib, L, world, rank, device
        self.cap = list(capacity_rows)
        self.own, handles = [], []
        for nc, cap in zip(n_tables_cols, capacity_rows):
            ptrs, hs = [], []
            for _ in range(nc):
                p = C.c_void_p()
                L.check(lib.tq_device_alloc(cap * 8, C.byref(p)))
                h = (C.c_ubyte * 64)()
                L.check(lib.tq_ipc_get_handle(p, h))
                ptrs.append(p.value)
                hs.append(bytes(h))
            self.own.append(ptrs)
            handles.append(hs)
        gathered = [None] * world
        dist.all_gather_object(gathered, handles)
        self.peer_ptr = []   # [dst][table][col] -> address valid in THIS process
        self._opened = []
        for dst in range(world):
            if dst == rank:
                self.peer_ptr.append(self.own)
                continue
            tabs = []
            for hs in gathered[dst]:
                cols = []
                for hb in hs:
                    p = C.c_void_p()
                    buf = (C.c_ubyte * 64).from_buffer_copy(hb)
                    L.check(lib.tq_ipc_open_handle(buf, C.byref(p)))
                    cols.append(p.value)
                    self._opened.append(p.value)
                tabs.append(cols)
            self.peer_ptr.append(tabs)
        self.recv = [[RawCol(p, cap) for p in ptrs] for ptrs, cap in zip(self.own, self.cap)]

    def close(self):
        for p in self._opened:
            self.lib.tq_ipc_close_handle(C.c_void_p(p))
        self._opened = []
        dist.barrier()  # nobody frees a buffer a peer still has mapped
        for ptrs in self.own:
            for p in ptrs:
                self.lib.tq_device_free(C.c_void_p(p))
        self.own = []

    def counts(self, key_cols):
        """local rows per destination for each table: tq_partition_count_device"""
        out = []
        for k in key_cols:
            n = int(k.numel())
            c = (C.c_int64 * self.world)()
            col = _tq_cols(self.L, [k], n)
            self.L.check(self.lib.tq_partition_count_device(col, n, self.world, c))
            out.append(list(c))
        return out

    def plan(self, local_counts):
        """all-gather the count matrices; returns (write offsets [table][dst], rows arriving here [table])"""
        mine = torch.tensor(local_counts, dtype=torch.int64, device=self.dev)      # [tables][world]
        allc = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allc, mine)
        m = torch.stack(allc).cpu().numpy()                                      # [src][table][dst]
        offs = [[int(m[: self.rank, t, d].sum()) for d in range(self.world)] for t in range(m.shape[1])]
        arriving = [int(m[:, t, self.rank].sum()) for t in range(m.shape[1])]
        for t, a in enumerate(arriving):
            if a > self.cap[t]:
                raise RuntimeError(f"receive buffer of table {t} too small: {a} rows arriving, capacity {self.cap[t]} (skewed keys)")
        return offs, arriving

    def push(self, table, cols, offs, async_op=False):
        n = int(cols[0].numel())
        ncols = len(cols)
        dest = (C.c_void_p * (self.world * ncols))()
        for d in range(self.world):
            for c in range(ncols):
                dest[d * ncols + c] = self.peer_ptr[d][table][c]
        o = (C.c_int64 * self.world)(*offs)
        fn = self.lib.tq_partition_push_device_async if async_op else self.lib.tq_partition_push_device
        self.L.check(fn(ncols, _tq_cols(self.L, cols, n), 0, n, self.world, dest, o))

    def wait(self):
        self.L.check(self.lib.tq_partition_push_wait())


def distributed_join(build_cols, probe_cols, world, rank, partition_fn, local_join_fn, group=None):
    """build_cols / probe_cols: lists of tensors holding this rank's row shard; key = column 0 of each side.
    partition_fn(cols, world) -> (partitioned cols, offsets[world+1]);  local_join_fn(build, probe) -> result."""
    b_part, b_off = partition_fn(build_cols, world)
    b_recv, _ = exchange(b_part, b_off, world, rank, group)
    p_part, p_off = partition_fn(probe_cols, world)
    p_recv, _ = exchange(p_part, p_off, world, rank, group)
    return local_join_fn(b_recv, p_recv)


def distributed_agg(cols, world, rank, partial_fn, partition_fn, final_fn, group=None):
    """The reference's partial -> shuffle -> final HashAgg (executor/aggregate.go:96-133,352-356,424-457) across ranks:
    partial_fn(cols) -> partial rows as a list of int64 tensors, GROUP BY key first (one row per LOCAL group:
    rows/world -> <= NDV rows, so the exchange moves groups, not input rows); partition_fn as in distributed_join
    (same key -> rank rule); final_fn(received partial rows) -> this rank's final groups (MergePartialResult)."""
    partials = partial_fn(cols)
    part, off = partition_fn(partials, world)
    recv, _ = exchange(part, off, world, rank, group)
    return final_fn(recv)


# ---------------------------------------------------------------------------------------------- GPU plumbing
def gpu_agg_fns(lib, L, types, group_col, funcs, est_groups=0):
    """(partial_fn, final_fn) for distributed_agg on device-resident NOT NULL int64/float64 columns: Partial1 handle ->
    tq_agg_export_partial; Final handle <- tq_agg_merge_partial.  Columns are declared TQ_TYPE_NOT_NULL: the partial
    states then carry no NULLs, so the exchange moves plain 8-byte columns (nullable inputs: merge locally instead)."""
    def make():
        it = (C.c_int32 * len(types))(*[t | L.TQ_TYPE_NOT_NULL for t in types])
        gb = (C.c_int32 * 1)(group_col)
        fa = (L.TQAggFunc * len(funcs))(*[L.TQAggFunc(f, a) for f, a in funcs])
        d = L.TQAggDesc(len(types), it, 1, gb, len(funcs), fa, est_groups)
        h = C.c_void_p()
        L.check(lib.tq_agg_create(C.byref(d), C.byref(h)))
        return h, (it, gb, fa)

    def partial_fn(cols):
        torch.cuda.synchronize()   # the library works on its own stream
        h, keep = make()
        try:
            n = int(cols[0].numel())
            if n:
                L.check(lib.tq_agg_put(h, _tq_cols(L, cols, n), L.TQ_MEM_DEVICE))
            width = C.c_int32(0)
            L.check(lib.tq_agg_partial_width(h, C.byref(width)))
            out = (L.TQColumn * width.value)()
            rows = C.c_int64(0)
            L.check(lib.tq_agg_export_partial(h, out, C.byref(rows)))
            # the lent arrays die with the handle: copy them into tensors the exchange can own
            res = []
            for c in range(width.value):
                t = torch.empty(rows.value, dtype=torch.int64, device=cols[0].device)
                if rows.value:
                    torch.cuda.current_stream().synchronize()
                    L.check(lib.tq_memcpy_d2d(t.data_ptr(), out[c].data, rows.value * 8))
                res.append(t)
        finally:
            lib.tq_agg_destroy(h)
        return res

    def final_fn(recv):
        torch.cuda.synchronize()
        h, keep = make()
        try:
            n = int(recv[0].numel())
            if n:
                L.check(lib.tq_agg_merge_partial(h, _tq_cols(L, recv, n), L.TQ_MEM_DEVICE))
            L.check(lib.tq_agg_eof(h))
            out = (L.TQColumn * len(funcs))()
            rows, eof = C.c_int64(0), C.c_int32(0)
            L.check(lib.tq_agg_next_device(h, out, C.byref(rows), C.byref(eof)))
            from .chunk import device_to_host
            res = []
            for i in range(len(funcs)):
                t = C.c_int32(0)
                L.check(lib.tq_agg_output_type(h, i, C.byref(t)))
                res.append(device_to_host(t.value, out[i].data, out[i].null_bitmap, rows.value))
        finally:
            lib.tq_agg_destroy(h)
        return res
    return partial_fn, final_fn


def _tq_cols(L, tensors, n):
    arr = (L.TQColumn * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i].length, arr[i].data, arr[i].null_bitmap, arr[i].offsets = n, t.data_ptr(), None, None
    return arr


def gpu_partition_fn(lib, L):
    def fn(cols, world):
        n = int(cols[0].numel())
        out = [torch.empty_like(c) for c in cols]
        offs = (C.c_int64 * (world + 1))()
        types = (C.c_int32 * len(cols))(*([1] * len(cols)))
        torch.cuda.synchronize()
        L.check(lib.tq_partition_device(len(cols), _tq_cols(L, cols, n), types, 0, n, world, _tq_cols(L, out, n), offs))
        return out, list(offs)
    return fn


def join_begin(lib, L, build, n_probe_cols=2):
    """create the join and build its table from device-resident build columns (key = column 0)"""
    nb = int(build[0].numel())
    t_b = (C.c_int32 * len(build))(*([1] * len(build)))
    t_p = (C.c_int32 * n_probe_cols)(*([1] * n_probe_cols))
    k = (C.c_int32 * 1)(0)
    d = L.TQJoinDesc(0, 1, len(build), t_b, n_probe_cols, t_p, 1, k, k, 0)
    h = C.c_void_p()
    L.check(lib.tq_join_create(C.byref(d), C.byref(h)))
    try:
        if nb:
            L.check(lib.tq_join_put_build(h, _tq_cols(L, build, nb), L.TQ_MEM_DEVICE))
        L.check(lib.tq_join_finalize_build(h))
    except Exception:
        lib.tq_join_destroy(h)
        raise
    return (h, len(build))


def join_finish(lib, L, handle, probe, keep_result=False):
    """probe with device-resident columns, drain, destroy.  Returns (rows, stats[, columns])."""
    h, n_build_cols = handle
    npr = int(probe[0].numel())
    total = 0
    result = None
    try:
        if npr:
            L.check(lib.tq_join_put_probe(h, _tq_cols(L, probe, npr), None, L.TQ_MEM_DEVICE))
        L.check(lib.tq_join_probe_eof(h))
        out = (L.TQColumn * (n_build_cols + len(probe)))()
        n, eof = C.c_int64(0), C.c_int32(0)
        while True:
            L.check(lib.tq_join_next_device(h, out, C.byref(n), C.byref(eof)))
            if n.value == 0 and eof.value:
                break
            total += n.value
            if keep_result and n.value:
                from .chunk import device_to_host
                result = [device_to_host(1, out[c].data, None, n.value) for c in range(n_build_cols + len(probe))]
        st = (C.c_int64 * 8)()
        lib.tq_join_stats(h, st)
    finally:
        lib.tq_join_destroy(h)
    return (total, list(st), result) if keep_result else (total, list(st))


def gpu_local_join(lib, L, build, probe, keep_result=False):
    """inner join, key = column 0 of both sides, int64 columns, inputs resident in HBM."""
    return join_finish(lib, L, join_begin(lib, L, build, len(probe)), probe, keep_result)


def bench_distributed_join(args, rank, world, local_rank, dist_mod, peak, peak_src):
    """bench.py N>1: weak scaling — every rank holds build_rows x probe_rows of a world-times larger join."""
    import statistics

    from . import _lib as L
    lib = L.load()
    dev = torch.device("cuda", local_rank)
    n_b, n_p = args.build_rows, args.probe_rows
    N_b = n_b * world
    rng = np.random.default_rng(1000 + rank)
    # this rank's build shard: the keys k with k % world == rank, shuffled (a disjoint cover of [0, N_b)); B.v = 7k+1
    bk_h = rng.permutation(n_b).astype(np.int64) * world + rank
    pk_h = rng.integers(0, N_b, n_p, dtype=np.int64)
    bk = torch.from_numpy(bk_h).to(dev)
    bv = bk * 7 + 1
    pk = torch.from_numpy(pk_h).to(dev)
    pv = torch.arange(n_p, dtype=torch.int64, device=dev) + rank * n_p
    part = gpu_partition_fn(lib, L)
    mode = os.environ.get("TQ_DIST_EXCHANGE", "push")
    use_peer = mode == "peer"
    px = None
    pushx = None
    if mode == "push":
        try:
            slack = lambda n: int(n * 1.25) + (1 << 16)   # keys are hash-spread: +25 % covers the imbalance
            pushx = PushExchange(lib, L, world, rank, dev, [2, 2], [slack(n_b), slack(n_p)])
        except Exception as e:
            if rank == 0:
                print(f"[dist] push exchange unavailable ({e}); falling back", flush=True)
            pushx = None
        okp = torch.tensor([1 if pushx is not None else 0], device=dev)
        dist_mod.all_reduce(okp, op=dist_mod.ReduceOp.MIN)
        if int(okp) == 0:
            pushx, use_peer = None, True
    if use_peer:
        try:
            px = PeerExchange(world, rank, dev, [2, 2], [n_b, n_p])
        except Exception as e:  # no CUDA IPC in this container: NCCL all_to_all instead
            if rank == 0:
                print(f"[dist] peer-memory exchange unavailable ({e}); using NCCL all_to_all", flush=True)
            px = None
    ok = torch.tensor([1 if px is not None else 0], device=dev)
    dist_mod.all_reduce(ok, op=dist_mod.ReduceOp.MIN)
    if int(ok) == 0:
        px = None

    phase_ms = {}

    def tick(name, t0):
        torch.cuda.synchronize()
        lib.tq_device_synchronize()
        phase_ms[name] = phase_ms.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return time.perf_counter()

    profile_phases = os.environ.get("TQ_DIST_PHASES") == "1"  # extra syncs per phase (diagnostics only; slows the step)

    def part_into(cols, outs):
        n = int(cols[0].numel())
        offs = (C.c_int64 * (world + 1))()
        types = (C.c_int32 * len(cols))(*([1] * len(cols)))
        L.check(lib.tq_partition_device(len(cols), _tq_cols(L, cols, n), types, 0, n, world, _tq_cols(L, outs, n), offs))
        return list(offs)

    def step_peer():
        t = time.perf_counter()
        b_off = part_into([bk, bv], px.send[0])   # tq_partition_device synchronises its stream before returning
        p_off = part_into([pk, pv], px.send[1])
        if profile_phases:
            t = tick("partition", t)
        offs = px.offsets_matrix([b_off, p_off])  # also orders every rank's scatter before anyone pulls
        b_recv, _, ev_b = px.pull(0, offs[:, 0, :])
        p_recv, _, ev_p = px.pull(1, offs[:, 1, :])
        for ev in ev_b:
            ev.synchronize()
        if profile_phases:
            t = tick("pull_build(+probe in flight)", t)
        h = join_begin(lib, L, b_recv)            # the probe rows are still arriving while the table is built
        for ev in ev_p:
            ev.synchronize()
        if profile_phases:
            t = tick("build+pull_probe", t)
        rows, st = join_finish(lib, L, h, p_recv)
        if profile_phases:
            t = tick("local_probe", t)
        dist_mod.barrier()                        # nobody overwrites its send buffers until every peer has pulled
        return rows, st, int(p_recv[0].numel())

    def step_nccl():
        t = time.perf_counter()
        b_part, b_off = part([bk, bv], world)
        p_part, p_off = part([pk, pv], world)
        if profile_phases:
            t = tick("partition", t)
        torch.cuda.synchronize()
        b_cnt, p_cnt = exchange_counts([b_off, p_off], world, rank, dev)
        b_recv, _ = exchange(b_part, b_off, world, rank, recv_counts=b_cnt)
        # the probe rows travel while the hash table is built from the build rows that already arrived
        p_recv, _, works = exchange(p_part, p_off, world, rank, recv_counts=p_cnt, async_op=True)
        torch.cuda.current_stream().synchronize()
        if profile_phases:
            t = tick("exchange_build", t)
        h = join_begin(lib, L, b_recv)
        for w in works:
            w.wait()
        torch.cuda.synchronize()
        if profile_phases:
            t = tick("build+exchange_probe", t)
        rows, st = join_finish(lib, L, h, p_recv)
        if profile_phases:
            t = tick("local_probe", t)
        return rows, st, int(p_recv[0].numel())

    n_chunks = max(1, int(os.environ.get("TQ_DIST_CHUNKS", "1")))
    bounds = [n_p * i // n_chunks for i in range(n_chunks + 1)]
    pk_c = [pk[bounds[i]:bounds[i + 1]] for i in range(n_chunks)]
    pv_c = [pv[bounds[i]:bounds[i + 1]] for i in range(n_chunks)]

    def step_push():
        """count -> one all-gather -> push build -> [push probe chunk i+1 over NVLink || probe chunk i locally]"""
        t = time.perf_counter()
        cnt = pushx.counts([bk] + pk_c)                      # rows per destination: build, then every probe chunk
        mine = torch.tensor(cnt, dtype=torch.int64, device=dev)
        allc = [torch.empty_like(mine) for _ in range(world)]
        dist_mod.all_gather(allc, mine)                      # also: every rank has finished consuming the previous step
        m = torch.stack(allc).cpu().numpy()                  # [src][table][dst]; table 0 = build, 1.. = probe chunks
        arr_b = int(m[:, 0, rank].sum())
        arr_c = [int(m[:, 1 + i, rank].sum()) for i in range(n_chunks)]
        if arr_b > pushx.cap[0] or sum(arr_c) > pushx.cap[1]:
            raise RuntimeError("receive buffer too small (skewed keys)")
        # write offsets: chunk i of the probe side lands behind chunks < i in the destination's receive buffer
        off_b = [int(m[:rank, 0, d].sum()) for d in range(world)]
        base = [[int(m[:, 1:1 + i, d].sum()) for d in range(world)] for i in range(n_chunks)]
        off_c = [[base[i][d] + int(m[:rank, 1 + i, d].sum()) for d in range(world)] for i in range(n_chunks)]
        if profile_phases:
            t = tick("count+plan", t)
        pushx.push(0, [bk, bv], off_b)                       # rows cross NVLink as the scatter kernel stores them
        dist_mod.barrier()                                   # every rank's build rows have landed
        pushx.push(1, [pk_c[0], pv_c[0]], off_c[0], async_op=True)   # first probe chunk travels while the table is built
        h = join_begin(lib, L, [x[:arr_b] for x in pushx.recv[0]])
        if profile_phases:
            t = tick("push_build+build", t)
        hh, nbc = h
        my_base = 0
        for i in range(n_chunks):
            pushx.wait()
            dist_mod.barrier()                               # chunk i has landed everywhere
            if i + 1 < n_chunks:
                pushx.push(1, [pk_c[i + 1], pv_c[i + 1]], off_c[i + 1], async_op=True)
            if arr_c[i]:
                cols = [RawCol(x.ptr + my_base * 8, arr_c[i]) for x in pushx.recv[1]]
                L.check(lib.tq_join_put_probe(hh, _tq_cols(L, cols, arr_c[i]), None, L.TQ_MEM_DEVICE))   # asynchronous: kernels queued
            my_base += arr_c[i]
        if profile_phases:
            t = tick("push_probe||local_probe", t)
        rows, st = join_finish(lib, L, h, [RawCol(0, 0), RawCol(0, 0)])
        if profile_phases:
            t = tick("drain", t)
        return rows, st, sum(arr_c)

    step = step_push if pushx is not None else (step_peer if px is not None else step_nccl)

    for _ in range(args.warmup):
        step()
    phase_ms.clear()
    from bench import ClockSampler
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    dist_mod.barrier()
    launches1 = lib.tq_kernel_launch_count()
    L.check(lib.tq_timer_start())
    t0 = time.perf_counter()
    rows_total, probe_ns = 0, []
    for _ in range(args.steps):
        rows, st, _ = step()
        rows_total += rows
        probe_ns.append(st[5])
    ms = C.c_float(0)
    L.check(lib.tq_timer_stop(C.byref(ms)))
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dist_mod.barrier()
    clocks = sampler.stop()
    launches2 = lib.tq_kernel_launch_count()
    t = torch.tensor([ms.value, float(rows_total), wall_ms], dtype=torch.float64, device=dev)
    tmax = t.clone()
    dist_mod.all_reduce(tmax, op=dist_mod.ReduceOp.MAX)
    tsum = t.clone()
    dist_mod.all_reduce(tsum, op=dist_mod.ReduceOp.SUM)
    ms_per_step = float(tmax[0]) / args.steps
    joined_per_step = float(tsum[1]) / args.steps
    value = joined_per_step / (ms_per_step * 1e-3)
    probe_s = statistics.mean(probe_ns) * 1e-9
    achieved = 64.0 * (joined_per_step / world) / probe_s / 1e9
    if rank != 0:
        return None
    return {
        "metric": "joined rows/sec on 1e8-row int64 equi-join", "value": value, "unit": "joined rows/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic",
        "config": {"workload": f"C5-style: int64 equi-join radix-partitioned over {world} GPUs; per GPU build={n_b} probe={n_p} (global {N_b} x {n_p * world}), "
                               "uniform keys, 100% match; partition -> grouped NCCL send/recv -> local join",
                   "build_rows_per_gpu": n_b, "probe_rows_per_gpu": n_p, "parallelism": f"key-hash partitions over {world} ranks", "exchange": ("push scatter into peer receive buffers (CUDA IPC, stores over NVLink)" if pushx is not None else
                                "peer-memory pull (CUDA IPC + NVLink copies)" if px is not None else "NCCL all_to_all_single"),
                   "l2": "inputs and outputs exceed the 126 MB L2; no flush needed"},
        "roofline": {"bound": "hbm", "kernel": "local probe pipeline (rank 0)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_row": 64, "kernel_ms": probe_s * 1e3},
        "e2e": {"value": value, "unit": "joined rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": "multi-GPU line: shards are generated in HBM; the host-buffer e2e figure is reported on the 1-GPU line"},
        "gpu_launches": int(launches2 - launches1), "clocks": clocks, "wall_ms_per_step_max": float(tmax[2]) / args.steps,
        "phase_ms_rank0": {k: v / args.steps for k, v in phase_ms.items()},
    }
