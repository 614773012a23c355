#!/usr/bin/env python
"""Kernel-vs-kernel diff of the Y-drop sweep: run the same gapped stage with the one-warp kernel
(LZB_DP_MODE=0) and the shared-memory kernel (LZB_DP_MODE=2), both dumping every row's
{LY, colEnd, best, used} (LZB_DP_DEBUG), and report the first row where they part.  GPU only."""
import os
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lastz_b200 import Engine, default_scoring, parse_seed, read_fasta, revcomp  # noqa: E402


def load(path):
    out = {}
    b = open(path, "rb").read()
    o = 0
    while o < len(b):
        magic, anchor, side, rows, status, cells, mode, nr = struct.unpack_from("<8I", b, o)
        assert magic == 0x44504447
        o += 32
        out[(anchor, side)] = dict(rows=rows, status=status, cells=cells, mode=mode,
                                   rec=np.frombuffer(b, dtype="<u4", count=nr * 4, offset=o).reshape(nr, 4))
        o += nr * 16
    return out


def run(tseq, qseq, strand, kw, label):
    seed = parse_seed()
    orc = Engine.oracle(); orc.set_scoring(default_scoring())
    to, qo = orc.build_seed_position_table(tseq, seed), orc.load_query(qseq)
    hsps, _ = orc.seed_hit_search(to, qo, seed, strand_id=strand)
    ao = orc.reduce_to_points(to, qo, hsps.copy())
    _, so, _ = orc.gapped_extend(to, qo, tseq, qseq, ao.copy(), **{k: v for k, v in kw.items() if k != "speculation"})
    dumps = {}
    for mode in (0, 1, 2):
        path = f"/tmp/dpdbg.{label}.{mode}.bin"
        if os.path.exists(path):
            os.remove(path)
        os.environ["LZB_DP_MODE"] = str(mode)
        os.environ["LZB_DP_DEBUG"] = path
        prod = Engine.product(0); prod.set_scoring(default_scoring())
        tp, qp = prod.build_seed_position_table(tseq, seed), prod.load_query(qseq)
        _, sp, _ = prod.gapped_extend(tp, qp, tseq, qseq, ao.copy(), **kw)
        prod.close()
        dumps[mode] = load(path)
        print(f"[{label}] mode {mode}: dpCells={sp.dpCells} rows={sp.dpRows} truncated={sp.truncated} extended={sp.anchorsExtended}"
              f"  (oracle cells={so.dpCells} truncated={so.truncated} extended={so.anchorsExtended})")
    for first in (0, 1):
        diff(dumps[first], dumps[2], f"mode {first}")


def diff(a, b, what):
    shown = 0
    for key in sorted(set(a) | set(b)):
        if key not in a or key not in b:
            print("  only in one dump:", key); continue
        x, y = a[key], b[key]
        n = min(len(x["rec"]), len(y["rec"]))
        d = np.nonzero((x["rec"][1:n] != y["rec"][1:n]).any(axis=1))[0]
        if x["rows"] != y["rows"] or x["cells"] != y["cells"] or len(d):
            print(f"  anchor {key[0]} side {key[1]}: warp rows={x['rows']} status={x['status']} cells={x['cells']} mode={x['mode']} | "
                  f"smem rows={y['rows']} status={y['status']} cells={y['cells']}")
            if len(d):
                r = int(d[0]) + 1
                for rr in range(max(1, r - 2), min(n, r + 4)):
                    print(f"    row {rr}: warp {x['rec'][rr].tolist()}  smem {y['rec'][rr].tolist()}")
            shown += 1
            if shown >= 4:
                break
    if not shown:
        print(f"  {what}: no per-row difference against the shared-memory kernel")


def main():
    golden = os.path.join(ROOT, "tests", "golden")
    tseq = read_fasta(os.path.join(golden, "pseudocat.fa"))[0][1]
    qseq = read_fasta(os.path.join(golden, "pseudopig.fa"))[0][1]
    run(tseq, qseq, 0, dict(speculation=1), "fixture+")
    run(tseq, revcomp(qseq), 3, dict(speculation=1), "fixture-")
    subprocess.check_call([os.path.join(ROOT, "tools", "gen_synth"), "300000", "20260925", "/tmp/t300.fa", "/tmp/q300.fa"])
    t, q = read_fasta("/tmp/t300.fa")[0][1], read_fasta("/tmp/q300.fa")[0][1]
    run(t, q, 0, dict(speculation=1, traceback_bytes=2 * 1024 * 1024), "synth2M")


if __name__ == "__main__":
    main()
