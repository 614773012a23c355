"""N>1 path on CPU: world_size 2 over gloo.  Each rank runs the hot path (oracle engine -- this is a
test) on its query interval; the segment tables are exchanged with the same all_gather bench.py uses
over NCCL; rank 0 checks the union against single-process runs on the same cuts, and against the
unmodified reference where oracle/_ref exists."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GEN_SYNTH, REF_CLI, ROOT

WORKER = r"""
import os, sys, pickle
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from lastz_b200 import Engine, default_scoring, parse_seed, read_fasta, revcomp
from lastz_b200.sharding import query_interval, gather_segment_tables, gather_to_rank0, pack_alignments
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
tseq = read_fasta(%(t)r)[0][1]; qseq = read_fasta(%(q)r)[0][1]
lo, hi = query_interval(len(qseq), rank, world)
eng = Engine.oracle(); eng.set_scoring(default_scoring()); seed = parse_seed()
T = eng.build_seed_position_table(tseq, seed)
tables, packed = [], []
for sid, s in ((0, qseq[lo:hi]), (3, revcomp(qseq[lo:hi]))):
    Q = eng.load_query(s)
    segs, st = eng.seed_hit_search(T, Q, seed, strand_id=sid)
    tables.append(segs)
    al, _, _ = eng.gapped_extend(T, Q, tseq, s, eng.reduce_to_points(T, Q, segs.copy()))
    packed.append(pack_alignments(al, sid))
mine = np.concatenate(tables)
parts = gather_segment_tables(mine, torch.device("cpu"))
als = gather_to_rank0(np.concatenate(packed).view(np.uint8), torch.device("cpu"))
assert (parts is None) == (rank != 0) and (als is None) == (rank != 0)      # a gather to rank 0, not an all_gather
if rank == 0:
    pickle.dump(([p.tolist() for p in parts], [a.tobytes() for a in als]), open(%(out)r, "wb"))
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_query_sharding(tmp_path):
    t, q = str(tmp_path / "t.fa"), str(tmp_path / "q.fa")
    subprocess.run([GEN_SYNTH, "120000", "20260925", t, q], check=True)
    out = str(tmp_path / "gathered.pkl")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, t=t, q=q, out=out))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    import pickle
    parts, als = pickle.load(open(out, "rb"))
    assert len(parts) == 2 and len(als) == 2

    # the same cuts in one process
    from lastz_b200 import Engine, default_scoring, parse_seed, read_fasta, revcomp
    from lastz_b200.sharding import query_interval, to_global, unpack_alignments
    tseq, qseq = read_fasta(t)[0][1], read_fasta(q)[0][1]
    eng = Engine.oracle()
    eng.set_scoring(default_scoring())
    seed = parse_seed()
    T = eng.build_seed_position_table(tseq, seed)
    for r in range(2):
        lo, hi = query_interval(len(qseq), r, 2)
        want, want_al = [], []
        for sid, s in ((0, qseq[lo:hi]), (3, revcomp(qseq[lo:hi]))):
            Q = eng.load_query(s)
            segs, _ = eng.seed_hit_search(T, Q, seed, strand_id=sid)
            want += segs.tolist()
            al, _, _ = eng.gapped_extend(T, Q, tseq, s, eng.reduce_to_points(T, Q, segs.copy()))
            want_al += [(sid, a["beg1"], a["beg2"], a["end1"], a["end2"], a["s"], a["ops"].tolist()) for a in al]
            # whole-query coordinates: the same bases on the whole strand as on the shard's strand
            whole = qseq if sid == 0 else revcomp(qseq)
            g = to_global(segs, lo, hi, len(qseq), revcomp=(sid != 0))
            for a, b in zip(segs[:50], g[:50]):
                assert whole[int(b["pos2"]):int(b["pos2"]) + int(b["length"])] == s[int(a["pos2"]):int(a["pos2"]) + int(a["length"])]
        assert parts[r] == want
        assert len(want) > 0
        got_al = unpack_alignments(np.frombuffer(als[r], dtype=np.uint32))
        assert [(a["strand"], a["beg1"], a["beg2"], a["end1"], a["end2"], a["s"], a["ops"].tolist()) for a in got_al] == want_al
        assert len(want_al) > 0

    # and the unmodified reference with the same cuts (q.fa[a..b]), forward strand coordinates
    if os.path.exists(REF_CLI):
        for r in range(2):
            lo, hi = query_interval(len(qseq), r, 2)
            ref = subprocess.run([REF_CLI, t, f"{q}[{lo + 1}..{hi}]", "--nogapped", "--format=segments"],
                                 capture_output=True, text=True, check=True).stdout.splitlines()[1:]
            ref_rows = [(int(x[1]), int(x[2]), int(x[4]), int(x[5]), x[6], int(x[7])) for x in (l.split() for l in ref)]
            dt = np.dtype([("hspId", "<u8"), ("pos1", "<u4"), ("pos2", "<u4"), ("length", "<u4"), ("s", "<i4"),
                           ("id", "<i4"), ("pad0", "<u4"), ("scoreCov", "<u8"), ("filter", "<i4"), ("pad1", "<u4")])
            got = np.array([tuple(x) for x in parts[r]], dtype=dt)
            fwd = got[got["id"] == 0]
            mine = [(int(g["pos1"]) + 1, int(g["pos1"] + g["length"]), int(g["pos2"]) + lo + 1,
                     int(g["pos2"] + g["length"]) + lo, "+", int(g["s"])) for g in fwd]
            assert mine == [x for x in ref_rows if x[4] == "+"]
