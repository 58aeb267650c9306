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


def assemble_grid(blocks):
    """blocks[r]: (n_qry, n_ref_r) block computed by rank r -> (n_qry, n_ref) in global reference order."""
    return torch.cat(list(blocks), dim=1)
