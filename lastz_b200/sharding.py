"""Query-interval sharding across ranks and the one exchange step.

The reference treats every query subrange `q.fa[a..b]` as an independent unit; rank r of N takes
query[r*L//N : (r+1)*L//N], runs the whole hot path on it against the replicated target, and the ranks' results --
48-byte segment records and the alignments with their edit scripts -- are gathered to rank 0: one all_gather of the
byte counts, then one gather of the payload (NCCL over NVLink on GPUs; gloo in the CPU tests).  Parity is defined
against the reference run with the same cuts: an HSP or alignment that would cross a cut ends at it.

Coordinates inside a shard are relative to the strand the shard was searched on, like the reference's subrange loads
(`startLoc` carries the offset, sequences.h:381ff); `to_global` turns them into coordinates of the whole query strand.
"""
import numpy as np


def query_interval(length, rank, world):
    """[lo, hi) of the query owned by `rank`."""
    return rank * length // world, (rank + 1) * length // world


def to_global(segs, lo, hi, query_len, revcomp=False):
    """Shard-relative pos2 -> position on the whole query strand.  The plus strand of shard [lo, hi) starts at lo; its
    reverse complement is the stretch [query_len - hi, query_len - lo) of the whole query's reverse complement."""
    out = segs.copy()
    out["pos2"] += np.uint32(query_len - hi if revcomp else lo)
    return out


def pack_alignments(aligns, strand_id):
    """alignments (dicts of Engine.gapped_extend) as one uint32 array: per alignment
    [strand, beg1, beg2, end1, end2, score, nops, ops...] -- the variable-length payload of the alignment gather."""
    words = []
    for a in aligns:
        words.append(np.array([strand_id, a["beg1"], a["beg2"], a["end1"], a["end2"], a["s"] & 0xFFFFFFFF, len(a["ops"])], dtype=np.uint32))
        words.append(np.asarray(a["ops"], dtype=np.uint32))
    return np.concatenate(words) if words else np.zeros(0, dtype=np.uint32)


def unpack_alignments(words):
    out, k = [], 0
    words = np.asarray(words, dtype=np.uint32)
    while k < len(words):
        strand, beg1, beg2, end1, end2, s, nops = (int(x) for x in words[k:k + 7])
        out.append(dict(strand=strand, beg1=beg1, beg2=beg2, end1=end1, end2=end2, s=s - (1 << 32) if s >= (1 << 31) else s,
                        ops=words[k + 7:k + 7 + nops].copy()))
        k += 7 + nops
    return out


def gather_to_rank0(raw, device):
    """Variable-length byte arrays (uint8) from every rank to rank 0: all_gather of the counts, then a gather of the
    payloads padded to the longest.  Returns the list of per-rank arrays on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    if world == 1:
        return [raw]
    rank = dist.get_rank()
    n = torch.tensor([raw.size], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    sizes = [int(c.item()) for c in counts]
    mx = max(max(sizes), 1)
    buf = torch.zeros(mx, dtype=torch.uint8, device=device)
    if raw.size:
        buf[:raw.size] = torch.from_numpy(raw).to(device)
    out = [torch.empty(mx, dtype=torch.uint8, device=device) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    return [o[:s].cpu().numpy() for o, s in zip(out, sizes)]


def gather_segment_tables(table, device):
    """lzb_segment arrays of every rank, on rank 0 (None elsewhere)."""
    parts = gather_to_rank0(np.ascontiguousarray(table).view(np.uint8).reshape(-1), device)
    if parts is None:
        return None
    return [np.frombuffer(p.tobytes(), dtype=table.dtype) for p in parts]
