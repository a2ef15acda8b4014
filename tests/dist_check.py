"""torchrun entry: distributed (N-GPU) hash join checked against the single-process oracle on the gathered inputs.
Launched by tests/test_gpu_dist.py (skipped when fewer than 2 GPUs are visible)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    from tinysql_b200 import _lib as L
    from tinysql_b200 import dist as D
    lib = L.load()
    L.check(lib.tq_init(local))
    dev = torch.device("cuda", local)
    rng = np.random.default_rng(7 + rank)
    nb, npr = 400000 + 1000 * rank, 1500000 + 777 * rank
    bk = rng.permutation(nb).astype(np.int64) * world + rank          # globally unique build keys
    bv = bk * 3 + 1
    pk = rng.integers(0, nb * world + 5000, npr).astype(np.int64)     # some probe keys have no match
    pv = np.arange(npr, dtype=np.int64) + rank * 10_000_000
    t = lambda a: torch.from_numpy(a).to(dev)
    part = D.gpu_partition_fn(lib, L)

    def local_join(b, p):
        return D.gpu_local_join(lib, L, b, p, keep_result=True)
    rows, st, res = D.distributed_join([t(bk), t(bv)], [t(pk), t(pv)], world, rank, part, local_join)
    # the region exchange (fused scatter + NVLink push, device-side flags, segmented local join) must give the same rows;
    # two steps: the second reuses the regions under a new epoch
    rj = D.RegionJoin(lib, L, world, rank, 400000 + 1000 * (world - 1), 1500000 + 777 * (world - 1), n_chunks=3)
    for _ in range(2):
        rows2, _, res2 = rj.step(t(bk), t(bv), t(pk), t(pv), keep_result=True)
        out2 = np.stack(res2, axis=1) if rows2 else np.zeros((0, 4), np.int64)
        srt = lambda m: m[np.lexsort(m.T[::-1])]
        assert rows2 == rows and np.array_equal(srt(out2), srt(np.stack([c.values for c in res], axis=1))), "region exchange differs from the NCCL exchange"
    rj.close()
    out = np.stack([c.values for c in res], axis=1) if rows else np.zeros((0, 4), np.int64)
    if rows:
        assert np.all(D.dest_rank_np(out[:, 0], world) == rank), "a key landed on the wrong rank"
    # gather everything on rank 0 and compare with the oracle
    gathered = [None] * world
    dist.all_gather_object(gathered, (out, bk, bv, pk, pv))
    if rank == 0:
        import oracle_py as O
        from tinysql_b200.chunk import INT64, Column
        got = np.concatenate([g[0] for g in gathered])
        BK, BV = np.concatenate([g[1] for g in gathered]), np.concatenate([g[2] for g in gathered])
        PK, PV = np.concatenate([g[3] for g in gathered]), np.concatenate([g[4] for g in gathered])
        want = O.hash_join(0, True, [INT64, INT64], [Column(INT64, BK), Column(INT64, BV)], [INT64, INT64], [Column(INT64, PK), Column(INT64, PV)], [0], [0])
        want = np.stack([c.values for c in want.cols], axis=1)
        assert got.shape == want.shape, (got.shape, want.shape)
        key = lambda m: m[np.lexsort(m.T[::-1])]
        assert np.array_equal(key(got), key(want))
        print(f"DIST_CHECK_OK world={world} rows={got.shape[0]}")
    # ---- distributed GROUP BY: Partial1 on every GPU -> partition the partial rows by key -> exchange -> Final merge
    na = 600000 + 999 * rank
    ak = rng.integers(0, 50000, na).astype(np.int64)
    ax = np.floor(rng.random(na) * 1024) / 8                       # dyadic doubles: sums exact in any order
    av = rng.integers(-1000, 1000, na).astype(np.int64)
    funcs = [(5, 0), (0, -1), (1, 1), (2, 1), (1, 2), (3, 2), (4, 1)]  # FIRSTROW(k) COUNT(*) SUM(x) AVG(x) SUM(v) MAX(v) MIN(x)
    pfn, ffn = D.gpu_agg_fns(lib, L, [1, 3, 1], 0, funcs, est_groups=50000)
    cols = [t(ak), torch.from_numpy(ax).to(dev), t(av)]
    res = D.distributed_agg(cols, world, rank, pfn, part, ffn)
    mine = np.stack([c.raw() for c in res], axis=1) if res[0].length else np.zeros((0, len(funcs)), np.uint64)
    if res[0].length:
        assert np.all(D.dest_rank_np(res[0].values, world) == rank), "a group was finalised on the wrong rank"
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, ak, ax, av))
    if rank == 0:
        import oracle_py as O
        from tinysql_b200.chunk import FLOAT64, INT64, Column
        got = np.concatenate([g[0] for g in gathered])
        AK, AX, AV = (np.concatenate([g[i] for g in gathered]) for i in (1, 2, 3))
        rc, want = O.hash_agg([INT64, FLOAT64, INT64], [Column(INT64, AK), Column(FLOAT64, AX), Column(INT64, AV)], [0], funcs, 2)
        wantm = np.stack([c.raw() for c in want.cols], axis=1)
        key = lambda m: m[np.lexsort(m.T[::-1])]
        assert rc == 0 and got.shape == wantm.shape, (got.shape, wantm.shape)
        assert np.array_equal(key(got), key(wantm)), "distributed GROUP BY differs from the oracle"
        print(f"DIST_AGG_OK world={world} groups={got.shape[0]}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
