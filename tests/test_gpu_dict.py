"""GPU tests of the sharded dictionary build (mashgpu_dict_*) and of jobs over pre-encoded rows (mashgpu_dist_open_encoded) on
ONE device: the hash-range exchange that NCCL does between ranks (mash_b200/shard.py) is replayed here by slicing arrays, so
that the C-ABI steps are pinned even where only one GPU is available.  The multi-rank run is tests/test_gpu_multi.py."""
import numpy as np
import pytest

from fixtures import synth_sketches, dense_rank_rows

pytestmark = pytest.mark.gpu


def _encode_in_shards(gpu, H, N, s, world):
    """The steps of shard.sharded_dictionary with the collectives replaced by array slicing."""
    import torch
    from mash_b200.shard import DictOps, shard_bounds
    dev = torch.device("cuda", 0)
    ops = DictOps(gpu)
    shards = []
    for b0, b1 in shard_bounds(H.shape[0], world):
        h = torch.from_numpy(H[b0:b1].view(np.int64).copy()).to(dev)
        n = torch.from_numpy(N[b0:b1].astype(np.int32)).to(dev)
        keys, slots = ops.local_sort(h, n, s)
        shards.append((h, n, keys, slots))
    allk = np.sort(np.concatenate([k.cpu().numpy().view(np.uint64) for _, _, k, _ in shards]))
    splitters = np.array([allk[(allk.size * d) // world] for d in range(1, world)], np.uint64) if allk.size else np.zeros(world - 1, np.uint64)
    send = [ops.split(k, splitters) if world > 1 else [k.numel()] for _, _, k, _ in shards]
    # "all-to-all": range d receives its slice of every shard's sorted keys
    codes_back = [[None] * world for _ in range(world)]
    n_distinct = []
    for d in range(world):
        parts = []
        for r, (_, _, k, _) in enumerate(shards):
            o = sum(send[r][:d])
            parts.append(k[o:o + send[r][d]])
        recv = torch.cat(parts) if parts else torch.empty(0, dtype=torch.int64, device=dev)
        codes, nd = ops.rank(recv)
        n_distinct.append(nd)
        o = 0
        for r in range(world):
            codes_back[r][d] = codes[o:o + send[r][d]]
            o += send[r][d]
    base = [sum(n_distinct[:d]) for d in range(world)]
    rows, neff = [], []
    for r, (h, n, k, slots) in enumerate(shards):
        codes = torch.cat(codes_back[r]) if k.numel() else torch.empty(0, dtype=torch.int32, device=dev)
        rr, ne = ops.scatter(codes, slots, send[r], base, h, n, s)
        rows.append(rr); neff.append(ne)
    return torch.cat(rows), torch.cat(neff), sum(n_distinct)


@pytest.mark.parametrize("world,n,s,stride", [(1, 40, 100, 100), (3, 50, 64, 64), (4, 9, 30, 45), (2, 33, 50, 20)])
def test_dict_entry_points_give_dense_ranks(gpu, world, n, s, stride):
    H, N, L = synth_sketches(n, stride, seed=100 + n, n_families=3, ragged=True)
    rows, neff, nd = _encode_in_shards(gpu, H, N, s, world)
    want_rows, want_neff = dense_rank_rows(H, N, s)
    assert np.array_equal(rows.cpu().numpy().view(np.uint32), want_rows)
    assert np.array_equal(neff.cpu().numpy().astype(np.uint32), want_neff)
    assert nd == int(want_rows[want_rows != 0xFFFFFFFF].max()) + 1


@pytest.mark.parametrize("triangle", [False, True])
def test_encoded_job_matches_oracle(gpu, oracle, triangle):
    import torch
    n, s = 150, 400
    H, N, L = synth_sketches(n, s, seed=61, n_families=4, ragged=True)
    rows, neff, _ = _encode_in_shards(gpu, H, N, s, 3)
    lens = torch.from_numpy(L.astype(np.int64)).to(rows.device)
    ks = 4.0 ** 21
    want = oracle.compare_all(H, N, L, H, N, L, s, 21, ks)
    for b0, b1 in ((0, 150), (40, 97), (149, 150)):
        job = gpu.dist_open_encoded(rows.data_ptr(), neff.data_ptr(), lens.data_ptr(), n, b0, b1 - b0, sketch_size=s, k=21, kmer_space=ks,
                                    keepalive=(rows, neff, lens))
        try:
            job.set_triangle(triangle)
            for pf in (0, 1):
                job.set_prefilter(pf)
                res = job.run(0, n)
                w = want[:, b0:b1]
                mask = np.ones_like(w["numer"], bool)
                if triangle:
                    mask = (np.arange(b0, b1)[None, :] < np.arange(n)[:, None])
                    assert np.all(res["numer"][~mask] == 0) and np.all(res["denom"][~mask] == 0)      # not computed: zeros
                assert np.array_equal(res["numer"][mask], w["numer"][mask]) and np.array_equal(res["denom"][mask], w["denom"][mask])
                assert np.all(np.abs(res["distance"][mask] - w["distance"][mask]) <= 1e-12)
                big = w["pvalue"] > 1e-305
                assert np.all(np.abs(res["pvalue"][mask & big] - w["pvalue"][mask & big]) <= 1e-12 * w["pvalue"][mask & big])
        finally:
            job.close()


def test_pass_list_skips_closed_form_pairs(gpu, oracle):
    # filtered run over a grid that is mostly unrelated pairs: the probe kernel writes nothing for them
    H, N, L = synth_sketches(200, 300, seed=71, n_families=20)
    ks = 4.0 ** 21
    job = gpu.dist_open(H, N, L, sketch_size=300, k=21, kmer_space=ks, max_distance=0.2, max_pvalue=1.0)
    try:
        job.set_prefilter(1)
        n_pass, lst = job.run_list(0, 200, 200 * 200)
        n2, lst2 = job.run_list(0, 200, 200 * 200)         # the persistent list buffers are reused
    finally:
        job.close()
    want = oracle.compare_all(H, N, L, H, N, L, 300, 21, ks, max_distance=0.2, max_pvalue=1.0)
    flat = np.flatnonzero(want["pass"].ravel())
    assert n_pass == flat.size == n2 and np.array_equal(lst["index"], flat.astype(np.uint64)) and np.array_equal(lst2["index"], lst["index"])
    assert np.array_equal(lst["numer"], want["numer"].ravel()[flat]) and np.array_equal(lst["denom"], want["denom"].ravel()[flat])
    assert np.all(np.abs(lst["distance"] - want["distance"].ravel()[flat]) <= 1e-12)
