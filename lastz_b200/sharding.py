"""Query-interval sharding across ranks and the one exchange step.

The reference treats every query subrange `q.fa[a..b]` as an independent unit; rank r of N takes
query[r*L//N : (r+1)*L//N], runs the whole hot path on it against the replicated target, and the ranks'
48-byte segment records are gathered with a single all_gather (NCCL over NVLink on GPUs; gloo in the
CPU tests).  Coordinates in a shard are relative to the shard, like the reference's subrange loads
(`startLoc` carries the offset, sequences.h:381ff); `to_global` adds it back.
"""
import numpy as np


def query_interval(length, rank, world):
    """[lo, hi) of the query owned by `rank`."""
    return rank * length // world, (rank + 1) * length // world


def to_global(segs, lo, strand_len=None, revcomp=False):
    """Shift shard-relative pos2 to whole-query coordinates (forward strand), or tag minus-strand
    records with their shard offset (they stay in the shard's reverse-complement frame, as in the
    reference's segments output)."""
    out = segs.copy()
    if not revcomp:
        out["pos2"] += np.uint32(lo)
    return out


def gather_segment_tables(table, device):
    """all_gather of variable-length lzb_segment arrays.  Returns the list of per-rank arrays."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return [table]
    raw = np.ascontiguousarray(table).view(np.uint8).reshape(-1)
    n = torch.tensor([raw.size], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    sizes = [int(c.item()) for c in counts]
    mx = max(max(sizes), 1)
    buf = torch.zeros(mx, dtype=torch.uint8, device=device)
    if raw.size:
        buf[:raw.size] = torch.from_numpy(raw.copy()).to(device)
    out = [torch.empty(mx, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(out, buf)
    return [np.frombuffer(o[:s].cpu().numpy().tobytes(), dtype=table.dtype) for o, s in zip(out, sizes)]
