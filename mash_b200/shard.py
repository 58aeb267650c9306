"""Multi-GPU plumbing (torch.distributed; NCCL on GPUs, gloo in the CPU tests).

Both hot paths shard without any reduction:
  * sketching: units (files / records) are independent -> contiguous blocks of units per rank, no collective;
    sketches stay on the rank that made them (the host gathers them in input order to write a .msh).
  * dist all-vs-all: the REFERENCE axis is sharded per rank and stays resident; every rank owns the query tile made of
    its own sketches and broadcasts it to the others (the one exchange step).  Rank r then computes the
    [all queries] x [its references] block; the full query-major grid is the concatenation of the blocks along the
    reference axis (reference order of `mash dist`: CommandDistance.cpp:213-232).
Only tensor movement happens here; all arithmetic is in libmashgpu.so.
"""
import torch
import torch.distributed as td


def shard_bounds(n, world):
    """Contiguous block partition of n items over `world` ranks: [(begin, end)] per rank."""
    per = (n + world - 1) // world if world > 0 else n
    return [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]


def exchange_query_tiles(hashes, n_hashes, lengths, counts=None, group=None):
    """Every rank broadcasts its tile of sketches (rows of `hashes`, with n_hashes / lengths); returns the
    concatenation over ranks in rank order, i.e. the global sketch order of `shard_bounds`.

    hashes: (m_r, s) int64/uint64 bit patterns, n_hashes: (m_r,) int32, lengths: (m_r,) int64. Tiles may have different
    row counts per rank (`counts` = list of row counts; gathered when None)."""
    world = td.get_world_size(group)
    rank = td.get_rank(group)
    dev = hashes.device
    if counts is None:
        c = torch.tensor([hashes.shape[0]], dtype=torch.int64, device=dev)
        allc = [torch.zeros_like(c) for _ in range(world)]
        td.all_gather(allc, c, group=group)
        counts = [int(x.item()) for x in allc]
    s = hashes.shape[1]
    out_h, out_n, out_l = [], [], []
    for r in range(world):
        if r == rank:
            bh, bn, bl = hashes.contiguous(), n_hashes.contiguous(), lengths.contiguous()
        else:
            bh = torch.empty((counts[r], s), dtype=hashes.dtype, device=dev)
            bn = torch.empty((counts[r],), dtype=n_hashes.dtype, device=dev)
            bl = torch.empty((counts[r],), dtype=lengths.dtype, device=dev)
        src = td.get_global_rank(group, r) if group is not None else r
        td.broadcast(bh, src=src, group=group)
        td.broadcast(bn, src=src, group=group)
        td.broadcast(bl, src=src, group=group)
        out_h.append(bh); out_n.append(bn); out_l.append(bl)
    return torch.cat(out_h), torch.cat(out_n), torch.cat(out_l), counts


def screen_allreduce(job, group=None):
    """Multi-GPU screen: reads are sharded over ranks, every rank holds the same reference table.  Sums the hit counters
    over ranks in place (NCCL all-reduce on the job's device counter array) and folds every other rank's mixture bottom-s
    list into this rank's job, after which job.finish() returns the global result on every rank."""
    ptr, n_slots = job.counters()
    dev = torch.device("cuda", job.eng.device)

    class _Cai:        # expose the library's device buffer to torch without a copy (CUDA array interface)
        __cuda_array_interface__ = {"shape": (n_slots,), "typestr": "<i4", "data": (ptr, False), "version": 3}

    counters = torch.as_tensor(_Cai(), device=dev)
    td.all_reduce(counters, op=td.ReduceOp.SUM, group=group)
    s = job.p.sketch_size
    mine = job.mixture()
    padded = torch.full((s,), -1, dtype=torch.int64, device=dev)
    padded[:mine.size] = torch.from_numpy(mine.view("int64")).to(dev)
    count = torch.tensor([mine.size], dtype=torch.int64, device=dev)
    world = td.get_world_size(group)
    all_lists = [torch.empty_like(padded) for _ in range(world)]
    all_counts = [torch.empty_like(count) for _ in range(world)]
    td.all_gather(all_lists, padded, group=group)
    td.all_gather(all_counts, count, group=group)
    rank = td.get_rank(group)
    for r in range(world):
        if r != rank:
            m = int(all_counts[r].item())
            if m:
                job.merge_mixture(all_lists[r][:m].cpu().numpy().view("uint64"))
    torch.cuda.synchronize(dev)


def assemble_grid(blocks):
    """blocks[r]: (n_qry, n_ref_r) block computed by rank r -> (n_qry, n_ref) in global reference order."""
    return torch.cat(list(blocks), dim=1)
