"""N>1 host logic on CPU (gloo, world_size 2): reference-axis sharding + query-tile broadcast + grid assembly.
The per-rank block is computed by the oracle here (no GPU in this container); on the GPU box the same plumbing feeds
libmashgpu (bench.py --gpus N, tests/test_gpu_multi.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as td
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


def _worker(rank, world, port, n, s, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fixtures import synth_sketches
        from mash_b200.shard import shard_bounds, exchange_query_tiles
        from oracle.pyoracle import Oracle
        H, N, L = synth_sketches(n, s, seed=17, n_families=3, ragged=True)      # same global set on every rank
        b0, b1 = shard_bounds(n, world)[rank]
        hl = torch.from_numpy(H[b0:b1].view(np.int64).copy())
        nl = torch.from_numpy(N[b0:b1].astype(np.int32))
        ll = torch.from_numpy(L[b0:b1].astype(np.int64))
        qh, qn, ql, counts = exchange_query_tiles(hl, nl, ll)
        assert counts == [e - b for b, e in shard_bounds(n, world)]
        assert np.array_equal(qh.numpy().view(np.uint64), H) and np.array_equal(qn.numpy(), N.astype(np.int32))
        orc = Oracle()
        ks = 4.0 ** 21
        block = orc.compare_all(H[b0:b1], N[b0:b1], L[b0:b1], qh.numpy().view(np.uint64), qn.numpy().astype(np.uint32),
                                ql.numpy().astype(np.uint64), s, 21, ks)
        np.save(os.path.join(tmp, f"block{rank}.npy"), block["numer"])
        td.barrier()
    finally:
        td.destroy_process_group()


def test_reference_axis_sharding_world2(tmp_path, oracle):
    from fixtures import synth_sketches
    from mash_b200.shard import assemble_grid, shard_bounds
    n, s, world = 37, 200, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, s, str(tmp_path)), nprocs=world, join=True)
    H, N, L = synth_sketches(n, s, seed=17, n_families=3, ragged=True)
    want = oracle.compare_all(H, N, L, H, N, L, s, 21, 4.0 ** 21)["numer"]
    blocks = [torch.from_numpy(np.load(tmp_path / f"block{r}.npy").astype(np.int64)) for r in range(world)]
    got = assemble_grid(blocks).numpy()
    assert got.shape == (n, n) and np.array_equal(got, want.astype(np.int64))
    assert shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_bounds(0, 2) == [(0, 0), (0, 0)]


def _dict_worker(rank, world, port, n, s, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fixtures import synth_sketches, NumpyDictOps
        from mash_b200.shard import shard_bounds, sharded_dictionary
        H, N, L = synth_sketches(n, s, seed=29, n_families=3, ragged=True)
        b0, b1 = shard_bounds(n, world)[rank]
        hl = torch.from_numpy(H[b0:b1].view(np.int64).copy())
        nl = torch.from_numpy(N[b0:b1].astype(np.int32))
        ll = torch.from_numpy(L[b0:b1].astype(np.int64))
        rows, n_eff, lens, counts, stats = sharded_dictionary(NumpyDictOps(), hl, nl, ll, s, n_samples=64)
        assert counts == [e - b for b, e in shard_bounds(n, world)]
        np.savez(os.path.join(tmp, f"dict{rank}.npz"), rows=rows.numpy().view(np.uint32), n_eff=n_eff.numpy(), lens=lens.numpy(),
                 ranked=stats["keys_ranked"], local=stats["keys_sorted_locally"])
        td.barrier()
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 37), (3, 20), (3, 2)])
def test_sharded_dictionary_protocol(tmp_path, world, n):
    """shard.sharded_dictionary (sample sort over hash ranges) with a numpy stand-in for the device steps: every rank must end
    with the dense ranks of the WHOLE collection, although no rank ever saw more than its rows and its hash range."""
    from fixtures import synth_sketches, dense_rank_rows
    s = 60
    mp.spawn(_dict_worker, args=(world, _free_port(), n, s, str(tmp_path)), nprocs=world, join=True)
    H, N, L = synth_sketches(n, s, seed=29, n_families=3, ragged=True)
    want_rows, want_neff = dense_rank_rows(H, N, s)
    total_ranked = 0
    for r in range(world):
        got = np.load(tmp_path / f"dict{r}.npz")
        assert np.array_equal(got["rows"], want_rows), f"rank {r}"
        assert np.array_equal(got["n_eff"].astype(np.uint32), want_neff) and np.array_equal(got["lens"].astype(np.uint64), L)
        total_ranked += int(got["ranked"])
    assert total_ranked == int(want_neff.sum())          # every hash was ranked exactly once, on the rank that owns its range


def test_numa_binding_helper_without_a_gpu_is_a_no_op():
    # bench.py binds a rank to the CPUs of its GPU's NUMA node; without CUDA (or without visible topology) it must leave the affinity alone
    import os
    from mash_b200.shard import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before
