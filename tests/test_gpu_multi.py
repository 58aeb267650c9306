"""Multi-GPU paths with the real kernels (needs >= 2 GPUs; skipped otherwise; runs on min(device_count, 8) ranks): NCCL plumbing
of mash_b200/shard.py.
  * dist: reference axis sharded per rank; (a) query tiles broadcast as hashes, every rank builds its own dictionary, and
    (b) the sharded dictionary build (sample sort over hash ranges + all-gather of the encoded rows) with an encoded job;
    both grids, assembled along the reference axis == oracle grid
  * screen: reads sharded, counters all-reduced, mixture lists merged on the device == oracle on the whole stream
  * sketch: units sharded, no collective == oracle per unit
Run on the 2- and 8-GPU lease: `gpurun --gpus N -- python -m pytest tests/test_gpu_multi.py -m gpu -q` (logs: profiles/r02_multi_gpu_pytest_*.log)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as td
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import mash_b200
        from mash_b200.shard import shard_bounds, exchange_query_tiles, screen_allreduce, sharded_dictionary, DictOps, warm_collectives
        from fixtures import synth_sketches, synth_genome
        eng = mash_b200.Engine(rank)
        dev = torch.device("cuda", rank)
        # ---- dist
        n, s = 150, 400
        H, N, L = synth_sketches(n, s, seed=23, n_families=4, ragged=True)
        b0, b1 = shard_bounds(n, world)[rank]
        hl = torch.from_numpy(H[b0:b1].view(np.int64).copy()).to(dev)
        nl = torch.from_numpy(N[b0:b1].astype(np.int32)).to(dev)
        ll = torch.from_numpy(L[b0:b1].astype(np.int64)).to(dev)
        qh, qn, ql, _ = exchange_query_tiles(hl, nl, ll)
        torch.cuda.synchronize()
        ref = mash_b200._capi._Set(hl.data_ptr(), nl.data_ptr(), ll.data_ptr(), on_device=True, n=b1 - b0, stride=s)
        qry = mash_b200._capi._Set(qh.data_ptr(), qn.data_ptr(), ql.data_ptr(), on_device=True, n=n, stride=s)
        job = mash_b200._capi.DistJob(eng, ref, None, None, qry, None, None, s, 21, 4.0 ** 21, 1.0, 1.0)
        res = job.run(0, n)
        job.close()
        np.savez(os.path.join(tmp, f"dist{rank}.npz"), **res)
        # ---- dist (b): sharded dictionary, encoded rows all-gathered, lower triangle as well
        warm_collectives(dev)
        rows, n_eff, lens, counts, stats = sharded_dictionary(DictOps(eng), hl, nl, ll, s, n_samples=128)
        assert counts == [e - b for b, e in shard_bounds(n, world)] and stats["keys_sorted_locally"] == int(nl.clamp(max=s + 1).sum())
        ejob = eng.dist_open_encoded(rows.data_ptr(), n_eff.data_ptr(), lens.data_ptr(), n, b0, b1 - b0, sketch_size=s, k=21, kmer_space=4.0 ** 21,
                                     keepalive=(rows, n_eff, lens))
        eres = ejob.run(0, n)
        ejob.set_triangle(True)
        tres = ejob.run(0, n)
        ejob.close()
        np.savez(os.path.join(tmp, f"edist{rank}.npz"), **eres)
        np.savez(os.path.join(tmp, f"tdist{rank}.npz"), **tres)
        np.save(os.path.join(tmp, f"rows{rank}.npy"), rows.cpu().numpy().view(np.uint32))
        # ---- screen
        p = eng.params(k=21, s=300)
        g = [synth_genome(70 + i, 120_000) for i in range(3)]
        rng = np.random.Generator(np.random.PCG64(5))
        reads = []
        for _ in range(4000):
            gg = g[int(rng.integers(0, 2))]
            a = int(rng.integers(0, gg.size - 150))
            reads.append(bytes(gg[a:a + 150]))
        refs = np.load(os.path.join(tmp, "screen_refs.npy"))
        refs_n = np.load(os.path.join(tmp, "screen_refs_n.npy"))
        r0, r1 = shard_bounds(len(reads), world)[rank]
        sjob = eng.screen_open(refs, refs_n, p)
        mine = reads[r0:r1]
        for c0 in range(0, len(mine), 700):
            sjob.feed(b"".join(b"*" + r for r in mine[c0:c0 + 700]))
        screen_allreduce(sjob)
        out = sjob.finish()
        sjob.close()
        np.savez(os.path.join(tmp, f"screen{rank}.npz"), **{k: np.asarray(v) for k, v in out.items()})
        # ---- sketch: units sharded, no collective
        units = [bytes(synth_genome(300 + u, 80_000 + 1000 * u)) for u in range(7)]
        u0, u1 = shard_bounds(len(units), world)[rank]
        h, nn, length = eng.sketch(units[u0:u1], p)
        np.savez(os.path.join(tmp, f"sketch{rank}.npz"), h=h, n=nn, length=length)
        td.barrier()
    finally:
        td.destroy_process_group()


def test_two_gpu_paths(tmp_path, oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from fixtures import synth_sketches, synth_genome, dense_rank_rows
    from mash_b200.shard import shard_bounds
    world = min(torch.cuda.device_count(), 8)
    print(f"multi-GPU parity on {world} ranks")
    po = oracle.params(k=21)
    g = [synth_genome(70 + i, 120_000) for i in range(3)]
    refs = np.full((3, 300), np.uint64(2**64 - 1)); refs_n = np.zeros(3, np.uint32)
    for i, gg in enumerate(g):
        h, _, _ = oracle.sketch_unit([bytes(gg)], po, s=300)
        refs[i, :h.size] = h; refs_n[i] = h.size
    np.save(tmp_path / "screen_refs.npy", refs); np.save(tmp_path / "screen_refs_n.npy", refs_n)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    # dist
    n, s = 150, 400
    H, N, L = synth_sketches(n, s, seed=23, n_families=4, ragged=True)
    want = oracle.compare_all(H, N, L, H, N, L, s, 21, 4.0 ** 21)
    blocks = [np.load(tmp_path / f"dist{r}.npz") for r in range(world)]
    numer = np.concatenate([b["numer"] for b in blocks], axis=1)
    denom = np.concatenate([b["denom"] for b in blocks], axis=1)
    dist = np.concatenate([b["distance"] for b in blocks], axis=1)
    assert np.array_equal(numer, want["numer"]) and np.array_equal(denom, want["denom"])
    assert np.all(np.abs(dist - want["distance"]) <= 1e-12)
    # dist (b): every rank holds the dense ranks of the whole collection; encoded jobs give the same grid; triangle too
    want_rows, _ = dense_rank_rows(H, N, s)
    eb = [np.load(tmp_path / f"edist{r}.npz") for r in range(world)]
    tb = [np.load(tmp_path / f"tdist{r}.npz") for r in range(world)]
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"rows{r}.npy"), want_rows)
    for key in ("numer", "denom"):
        assert np.array_equal(np.concatenate([b[key] for b in eb], axis=1), want[key])
    assert np.array_equal(np.concatenate([b["distance"] for b in eb], axis=1), dist)
    pv = np.concatenate([b["pvalue"] for b in eb], axis=1)
    big = want["pvalue"] > 1e-305
    assert np.all(np.abs(pv[big] - want["pvalue"][big]) <= 1e-12 * want["pvalue"][big])
    lower = np.arange(n)[None, :] < np.arange(n)[:, None]
    tn = np.concatenate([b["numer"] for b in tb], axis=1)
    assert np.array_equal(tn[lower], want["numer"][lower]) and np.all(tn[~lower] == 0)
    # screen: every rank holds the global answer
    rng = np.random.Generator(np.random.PCG64(5))
    reads = []
    for _ in range(4000):
        gg = g[int(rng.integers(0, 2))]
        a = int(rng.integers(0, gg.size - 150))
        reads.append(bytes(gg[a:a + 150]))
    ws = oracle.screen(refs, refs_n, [b"".join(b"*" + r for r in reads)], po, s=300)
    for r in range(world):
        got = np.load(tmp_path / f"screen{r}.npz")
        assert np.array_equal(got["shared"], ws["shared"]) and np.array_equal(got["median"], ws["median"])
        assert int(got["set_size"]) == ws["set_size"] and np.array_equal(got["mixture"], ws["mixture"])
    # sketch
    units = [bytes(synth_genome(300 + u, 80_000 + 1000 * u)) for u in range(7)]
    for r in range(world):
        u0, u1 = shard_bounds(len(units), world)[r]
        got = np.load(tmp_path / f"sketch{r}.npz")
        for i, u in enumerate(range(u0, u1)):
            oh, _, olen = oracle.sketch_unit([units[u]], po, s=300)
            assert got["length"][i] == olen and np.array_equal(got["h"][i, :got["n"][i]], oh)
