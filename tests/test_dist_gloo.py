"""CPU suite, part 4: the N>1 host path (count exchange + grouped send/recv + local join) with world_size 2 and 3 over gloo.
The device kernels are replaced by injected CPU stand-ins (numpy partitioner using the same key -> rank rule, oracle join);
what is under test is tinysql_b200.dist.exchange / distributed_join."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_py as O
    from tinysql_b200 import dist as D
    from tinysql_b200.chunk import INT64, Column

    rng = np.random.default_rng(100 + rank)
    nb, npr = 3000 + 17 * rank, 20000 + 31 * rank  # ragged shards
    bk = rng.integers(0, 5000, nb)                  # duplicate keys across ranks
    bv = np.arange(nb) + 1000000 * rank
    pk = rng.integers(0, 6000, npr)
    pv = np.arange(npr) + 1000000 * rank

    def partition_fn(cols, world):
        d = D.dest_rank_np(cols[0].numpy(), world)
        order = np.argsort(d, kind="stable")
        counts = np.bincount(d, minlength=world)
        return [c[torch.from_numpy(order)] for c in cols], [0] + list(np.cumsum(counts))

    def local_join(b, p):
        bc = [Column(INT64, t.numpy()) for t in b]
        pc = [Column(INT64, t.numpy()) for t in p]
        return O.hash_join(0, True, [INT64, INT64], bc, [INT64, INT64], pc, [0], [0])

    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64))
    res = D.distributed_join([t(bk), t(bv)], [t(pk), t(pv)], world, rank, partition_fn, local_join)
    # every key that arrived here must belong to this rank
    if res.num_rows():
        assert np.all(D.dest_rank_np(res.cols[0].values, world) == rank)
    np.savez(os.path.join(outdir, f"r{rank}.npz"), out=np.stack([c.values for c in res.cols], axis=1) if res.num_rows() else np.zeros((0, 4), np.int64),
             bk=bk, bv=bv, pk=pk, pv=pv)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_join_matches_single_process_oracle(tmp_path, world):
    import oracle_py as O
    from tinysql_b200.chunk import INT64, Column
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(world)]
    got = np.concatenate([p["out"] for p in parts])
    bk, bv = np.concatenate([p["bk"] for p in parts]), np.concatenate([p["bv"] for p in parts])
    pk, pv = np.concatenate([p["pk"] for p in parts]), np.concatenate([p["pv"] for p in parts])
    want = O.hash_join(0, True, [INT64, INT64], [Column(INT64, bk), Column(INT64, bv)], [INT64, INT64], [Column(INT64, pk), Column(INT64, pv)], [0], [0])
    want = np.stack([c.values for c in want.cols], axis=1)
    assert got.shape == want.shape
    key = lambda m: m[np.lexsort(m.T[::-1])]
    assert np.array_equal(key(got), key(want))


def _agg_worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tinysql_b200 import dist as D
    rng = np.random.default_rng(200 + rank)
    n = 30000 + 13 * rank
    k = rng.integers(0, 700, n)
    v = rng.integers(-50, 50, n)

    def partial_fn(cols):  # CPU stand-in for Partial1 + tq_agg_export_partial: rows (key, count, sum)
        kk, vv = cols[0].numpy(), cols[1].numpy()
        keys, inv = np.unique(kk, return_inverse=True)
        cnt = np.bincount(inv, minlength=len(keys)).astype(np.int64)
        sm = np.bincount(inv, weights=vv, minlength=len(keys)).astype(np.int64)
        return [torch.from_numpy(keys.astype(np.int64)), torch.from_numpy(cnt), torch.from_numpy(sm)]

    def partition_fn(cols, world):
        d = D.dest_rank_np(cols[0].numpy(), world)
        order = np.argsort(d, kind="stable")
        counts = np.bincount(d, minlength=world)
        return [c[torch.from_numpy(order)] for c in cols], [0] + list(np.cumsum(counts))

    def final_fn(recv):    # MergePartialResult: counts add, sums add
        kk, cc, ss = (t.numpy() for t in recv)
        keys, inv = np.unique(kk, return_inverse=True)
        return keys, np.bincount(inv, weights=cc, minlength=len(keys)).astype(np.int64), np.bincount(inv, weights=ss, minlength=len(keys)).astype(np.int64)

    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64))
    keys, cnt, sm = D.distributed_agg([t(k), t(v)], world, rank, partial_fn, partition_fn, final_fn)
    assert np.all(D.dest_rank_np(keys, world) == rank)   # every group is finalised on exactly one rank
    np.savez(os.path.join(outdir, f"a{rank}.npz"), keys=keys, cnt=cnt, sm=sm, k=k, v=v)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_agg_partial_shuffle_final(tmp_path, world):
    """partial -> shuffle -> final across ranks == the oracle's HashAgg over all rows (aggregate.go:96-133,352-356)"""
    import oracle_py as O
    from tinysql_b200.chunk import INT64, Column
    port = _free_port()
    mp.spawn(_agg_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(tmp_path, f"a{r}.npz")) for r in range(world)]
    got = sorted(zip(np.concatenate([p["keys"] for p in parts]).tolist(), np.concatenate([p["cnt"] for p in parts]).tolist(),
                     np.concatenate([p["sm"] for p in parts]).tolist()))
    k, v = np.concatenate([p["k"] for p in parts]), np.concatenate([p["v"] for p in parts])
    rc, want = O.hash_agg([INT64, INT64], [Column(INT64, k), Column(INT64, v)], [0], [(5, 0), (0, -1), (1, 1)], 2)
    assert rc == 0 and got == sorted(want.rows())


def test_dest_rank_matches_device_rule():
    """the numpy stand-in uses the same key -> rank rule as partition.cu (mix64(key) >> 40) % n_parts"""
    from tinysql_b200 import dist as D
    k = np.array([0, 1, 2, 12345678901234, -1, -(1 << 63)], dtype=np.int64)
    m = D.mix64_np(k.view(np.uint64))
    # murmur3 fmix64 known answers: fmix64(0) = 0, fmix64(1) = 0xb456bcfc34c2cb2c
    assert int(m[0]) == 0 and int(m[1]) == 0xB456BCFC34C2CB2C
    assert np.all(D.dest_rank_np(k, 8) == ((m >> np.uint64(40)) % np.uint64(8)).astype(np.int64))
