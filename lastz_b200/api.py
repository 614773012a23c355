"""Host-side mirror of the reference's operator interface for the seed-and-extend path.

Function names and argument meaning follow the reference (lastz 1.04.58) so the parity tests read
like its own call sites:

    build_seed_position_table  pos_table.h:230        -> Engine.build_seed_position_table
    seed_hit_search            seed_search.h:265      -> Engine.seed_hit_search
    reduce_to_points           gapped_extend.h:151    -> Engine.reduce_to_points
    gapped_extend              gapped_extend.h:153    -> Engine.gapped_extend

An ``Engine`` wraps one library implementing include/lastz_b200.h: ``Engine.product(device)`` is the
CUDA library (fails loudly without it or without a GPU); ``Engine.oracle()`` is the CPU
restatement and may be constructed only by tests/, smoke() and bench.py's cpu_baseline leg.
"""
import ctypes as C
import os

import numpy as np

from . import capi

_HOST_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "liblzb_host.so")

SEED_12OF19 = "1110100110010101111"   # seeds.h:86
SEED_14OF22 = "1110101100110010101111"

RCF_FORWARD, RCF_REVCOMP = 0, 3       # sequences.h:345-350


class _ScoreSet(C.Structure):         # lzb_host.h lzb_scoreset
    _fields_ = [("sub", C.c_int32 * 65536), ("masked", C.c_int32 * 65536),
                ("gapOpen", C.c_int32), ("gapExtend", C.c_int32),
                ("gapOpenSet", C.c_int), ("gapExtendSet", C.c_int),
                ("hspThresholdSet", C.c_int), ("gappedThresholdSet", C.c_int), ("xDropSet", C.c_int),
                ("yDropSet", C.c_int), ("stepSet", C.c_int),
                ("hspThreshold", C.c_int32), ("gappedThreshold", C.c_int32), ("xDrop", C.c_int32),
                ("yDrop", C.c_int32), ("step", C.c_uint32)]


_host = None


def host_lib():
    global _host
    if _host is None:
        if not os.path.exists(_HOST_LIB):
            raise RuntimeError(f"{_HOST_LIB} is missing: run __graft_entry__.build()")
        _host = C.CDLL(_HOST_LIB)
        _host.lzb_seed_parse.argtypes = [C.POINTER(capi.Seed), C.c_char_p, C.c_int]
        _host.lzb_scores_default.argtypes = [C.POINTER(_ScoreSet)]
        _host.lzb_scores_read_file.argtypes = [C.POINTER(_ScoreSet), C.c_char_p]
        _host.lzb_reduce_to_chain.restype = C.c_int32
        _host.lzb_reduce_to_chain.argtypes = [C.POINTER(capi.Segment), C.POINTER(C.c_uint64), C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        _host.lzb_merge_segments.restype = None
        _host.lzb_merge_segments.argtypes = [C.POINTER(capi.Segment), C.POINTER(C.c_uint64)]
    return _host


def reduce_to_chain(segs, scoring, diag_penalty=0, anti_penalty=0):
    """reduce_to_chain chain.c:497 with the penalties of lastz.c:3687 (chainScale 100, lastz.c:511): the HSP table
    cut down to its best chain, ordered by pos1.  Host C (csrc/host/chain.c), as in the reference."""
    a = np.ascontiguousarray(segs).copy()
    n = C.c_uint64(len(a))
    if len(a):
        host_lib().lzb_reduce_to_chain(a.ctypes.data_as(C.POINTER(capi.Segment)), C.byref(n), diag_penalty, anti_penalty,
                                       100, scoring.sub[ord("A") * 256 + ord("A")])
    return a[:n.value]


def merge_segments(segs):
    """merge_segments segment.c:1527: HSPs that share positions on one diagonal become one (what the reference does to the
    table of the recoverable hit processor and to anchors read from a file).  Host C (csrc/host/merge.c)."""
    a = np.ascontiguousarray(segs).copy()
    n = C.c_uint64(len(a))
    if len(a):
        host_lib().lzb_merge_segments(a.ctypes.data_as(C.POINTER(capi.Segment)), C.byref(n))
    return a[:n.value]


def parse_seed(pattern=SEED_12OF19, with_trans=1):
    """parse_seeds_string + create_seed_structure (seeds.c:321, lastz.c:9700)."""
    s = capi.Seed()
    host_lib().lzb_seed_parse(C.byref(s), pattern.encode(), with_trans)
    return s


def default_scoring(score_file=None):
    """HOXD70 (dna_utilities.c:137-148) or a scoring file (dna_utilities.c:581-628)."""
    ss = _ScoreSet()
    if score_file:
        host_lib().lzb_scores_read_file(C.byref(ss), score_file.encode())
    else:
        host_lib().lzb_scores_default(C.byref(ss))
    return ss


def upper_nuc_to_bits():
    t = (C.c_int8 * 256)(*([-1] * 256))
    for ch, v in zip(b"ACGT", range(4)):
        t[ch] = v
    return t


_COMP = bytes.maketrans(b"ACGTRYMKBDHVNSWacgtrymkbdhvnsw", b"TGCAYRKMVHDBNSWtgcayrkmvhdbnsw")


def revcomp(seq: bytes) -> bytes:
    """rev_comp_sequence sequences.c:7511."""
    return seq.translate(_COMP)[::-1]


def read_fasta(path):
    """[(header, bytes)] -- bases kept verbatim (case = soft mask), like load_fasta_sequence."""
    out, hdr, chunks = [], None, []
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if hdr is not None:
                    out.append((hdr, b"".join(chunks)))
                hdr, chunks = line.rstrip(b"\r\n").decode(), []
            else:
                chunks.append(b"".join(line.split()))
    if hdr is not None or chunks:
        out.append((hdr or "", b"".join(chunks)))
    return out


def segments_to_array(ptr, n):
    """Copy an lzb_segment[n] into a structured numpy array."""
    dt = np.dtype([("hspId", "<u8"), ("pos1", "<u4"), ("pos2", "<u4"), ("length", "<u4"), ("s", "<i4"),
                   ("id", "<i4"), ("pad0", "<u4"), ("scoreCov", "<u8"), ("filter", "<i4"), ("pad1", "<u4")])
    if n == 0:
        return np.zeros(0, dtype=dt)
    buf = C.string_at(ptr, n * 48)
    return np.frombuffer(buf, dtype=dt).copy()


class Engine:
    """One library + one context (device, stream, scoring)."""

    def __init__(self, lib, device=0):
        self.lib = lib
        self.ctx = lib.lzb_open(device)
        if not self.ctx:
            raise RuntimeError(lib.lzb_last_error().decode())
        self.backend = lib.lzb_backend().decode()
        self.ctb = upper_nuc_to_bits()
        self.scoring = None

    @classmethod
    def product(cls, device=0):
        return cls(capi.load_product(), device)

    @classmethod
    def oracle(cls):
        return cls(capi.load_oracle(), 0)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("FAILURE: " + self.lib.lzb_last_error().decode())

    def close(self):
        if self.ctx:
            self.lib.lzb_close(self.ctx)
            self.ctx = None

    def launches(self):
        return int(self.lib.lzb_launch_count(self.ctx))

    def set_scoring(self, ss):
        self.scoring = ss
        self._check(self.lib.lzb_set_scoring(self.ctx, ss.sub, ss.masked, ss.gapOpen, ss.gapExtend))

    # ---- pos_table.h:230
    def build_seed_position_table(self, seq1: bytes, seed, step=1, start=0, end=0):
        t = self.lib.lzb_target_build(self.ctx, seq1, len(seq1), start, end, self.ctb, C.byref(seed), step)
        if not t:
            raise RuntimeError("FAILURE: " + self.lib.lzb_last_error().decode())
        return t

    def limit_position_table(self, t, limit):
        """limit_position_table pos_table.h:240 (maxChasm 0): drop the words that occur more than `limit` times"""
        self._check(self.lib.lzb_target_limit(t, limit))

    def free_position_table(self, t):
        self.lib.lzb_target_free(t)

    def export_index(self, t, word_bits):
        counts = np.zeros(1 << word_bits, dtype=np.uint32)
        n = self.lib.lzb_target_export_index(t, counts.ctypes.data_as(C.POINTER(C.c_uint32)), None)
        pos = np.zeros(max(n, 1), dtype=np.uint32)
        self.lib.lzb_target_export_index(t, counts.ctypes.data_as(C.POINTER(C.c_uint32)),
                                         pos.ctypes.data_as(C.POINTER(C.c_uint32)))
        return counts, pos[:n]

    def load_query(self, seq2: bytes):
        q = self.lib.lzb_query_load(self.ctx, seq2, len(seq2))
        if not q:
            raise RuntimeError("FAILURE: " + self.lib.lzb_last_error().decode())
        return q

    def free_query(self, q):
        self.lib.lzb_query_free(q)

    # ---- seed_search.h:265 (+ process_for_simple_hit, xdrop_extend_seed_hit, collect_hsps)
    def seed_hit_search(self, t, q, seed, *, start=0, end=0, gf_extend=1, x_drop=910,
                        hsp_threshold=3000, entropy=True, hash_bits=16, self_compare=False,
                        same_strand=False, strand_id=RCF_FORWARD, plain_hits=False, gf_mismatches=0, recover_seeds=False, twin_spans=(0, 0), seed_queue=0, extend_ctas=0, search_limit=0):
        p = capi.SeedParams(start, end, gf_extend, x_drop, hsp_threshold, int(entropy), hash_bits,
                            int(self_compare), int(same_strand), strand_id, int(plain_hits), gf_mismatches, int(recover_seeds),
                            int(twin_spans[0]), int(twin_spans[1]), int(seed_queue), int(search_limit), int(extend_ctas))
        segs = C.POINTER(capi.Segment)()
        n = C.c_uint64(0)
        st = capi.SeedStats()
        self._check(self.lib.lzb_seed_hit_search(self.ctx, t, q, C.byref(seed), self.ctb, C.byref(p),
                                                 C.byref(segs), C.byref(n), C.byref(st)))
        arr = segments_to_array(segs, n.value)
        self.lib.lzb_free(C.cast(segs, C.c_void_p))
        return arr, st

    # ---- gapped_extend.h:151
    def reduce_to_points(self, t, q, anchors):
        a = np.ascontiguousarray(anchors)
        self._check(self.lib.lzb_reduce_to_points(self.ctx, t, q, a.ctypes.data_as(C.POINTER(capi.Segment)), len(a)))
        return a

    # ---- gapped_extend.h:153
    def gapped_extend(self, t, q, seq1: bytes, seq2: bytes, anchors, *, y_drop=9400, trim_to_peak=True,
                      score_threshold=3000, all_bounds=False, inhibit_trivial=False, identity_check=False,
                      traceback_bytes=80 * 1024 * 1024, speculation=64, max_paired_bases=0, overly_paired_keep=False):
        a = np.ascontiguousarray(anchors)
        p = capi.GappedParams(y_drop, int(trim_to_peak), score_threshold, int(all_bounds), int(inhibit_trivial),
                              int(identity_check), traceback_bytes, speculation, int(overly_paired_keep), int(max_paired_bases))
        lst = C.POINTER(capi.Alignel)()
        st = capi.GappedStats()
        self._check(self.lib.lzb_gapped_extend(self.ctx, t, q, seq1, seq2, a.ctypes.data_as(C.POINTER(capi.Segment)),
                                               len(a), C.byref(p), C.byref(lst), C.byref(st)))
        out = []
        node = lst
        while node:
            al = node.contents
            ops = np.ctypeslib.as_array(C.cast(C.addressof(al.script.contents) + 12, C.POINTER(C.c_uint32)),
                                        shape=(al.script.contents.len,)).copy()
            out.append(dict(beg1=al.beg1, beg2=al.beg2, end1=al.end1, end2=al.end2, s=al.s,
                            isTrivial=al.isTrivial, ops=ops))
            node = al.next
        self.lib.lzb_free_align_list(lst)
        return out, st, a
