"""ctypes binding of include/lastz_b200.h.

The product library is ``lastz_b200/csrc/liblastz_b200.so`` (hand-written sm_100a kernels).  It is
loaded explicitly and there is no fallback: if it is missing, or no GPU is present when a context
is opened, the call fails loudly.  ``load_oracle()`` loads the CPU restatement under ``oracle/`` --
test infrastructure that only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline
leg may use.

Reference interfaces mirrored (lastz 1.04.58): build_seed_position_table pos_table.h:230,
seed_hit_search seed_search.h:265, reduce_to_points / gapped_extend gapped_extend.h:151-159,
structs segment.h:46-64 and edit_script.h:30-61.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
PRODUCT_LIB = os.path.join(_HERE, "csrc", "liblastz_b200.so")
ORACLE_LIB = os.path.join(ROOT, "oracle", "liblzb_oracle.so")

MAX_PARTS = 32
MAX_FLIPS = 32


class Segment(C.Structure):          # segment.h:46-64 (48 bytes)
    _fields_ = [("hspId", C.c_uint64), ("pos1", C.c_uint32), ("pos2", C.c_uint32),
                ("length", C.c_uint32), ("s", C.c_int32), ("id", C.c_int32), ("pad0", C.c_uint32),
                ("scoreCov", C.c_uint64), ("filter", C.c_int32), ("pad1", C.c_uint32)]


class EditScript(C.Structure):       # edit_script.h:55-61
    _fields_ = [("size", C.c_uint32), ("len", C.c_uint32), ("tailOp", C.c_uint32), ("op", C.c_uint32 * 1)]


class Alignel(C.Structure):          # edit_script.h:30-41 (64 bytes)
    pass


Alignel._fields_ = [("next", C.POINTER(Alignel)), ("isTrivial", C.c_int32),
                    ("beg1", C.c_uint32), ("beg2", C.c_uint32), ("end1", C.c_uint32), ("end2", C.c_uint32),
                    ("s", C.c_int32), ("seq1", C.c_void_p), ("seq2", C.c_void_p),
                    ("script", C.POINTER(EditScript)), ("hspId", C.c_uint64)]


class Seed(C.Structure):             # seeds.h:37-76 (flat)
    _fields_ = [("length", C.c_int32), ("weight", C.c_int32), ("numParts", C.c_int32),
                ("withTrans", C.c_int32), ("numFlips", C.c_int32),
                ("shift", C.c_int32 * MAX_PARTS), ("mask", C.c_uint32 * MAX_PARTS),
                ("transFlips", C.c_uint32 * MAX_FLIPS)]


class SeedParams(C.Structure):
    _fields_ = [("start", C.c_uint32), ("end", C.c_uint32), ("gfExtend", C.c_int32), ("xDrop", C.c_int32),
                ("hspThreshold", C.c_int32), ("entropy", C.c_int32), ("hashBits", C.c_int32),
                ("selfCompare", C.c_int32), ("sameStrand", C.c_int32), ("strandId", C.c_int32),
                ("plainHits", C.c_int32), ("gfMismatches", C.c_int32), ("recoverSeeds", C.c_int32),
                ("twinMinSpan", C.c_int32), ("twinMaxSpan", C.c_int32), ("seedQueueSize", C.c_int32), ("searchLimit", C.c_uint32), ("extendCtasPerSm", C.c_int32)]


class SeedStats(C.Structure):
    _fields_ = [("wordsInQuery", C.c_uint64), ("rawSeedHits", C.c_uint64), ("extensions", C.c_uint64),
                ("bpExtended", C.c_uint64), ("hsps", C.c_uint64), ("seconds", C.c_double),
                ("kernelSeconds", C.c_double * 12), ("kernelLaunches", C.c_uint64 * 12)]


class GappedParams(C.Structure):
    _fields_ = [("yDrop", C.c_int32), ("trimToPeak", C.c_int32), ("scoreThreshold", C.c_int32),
                ("allBounds", C.c_int32), ("inhibitTrivial", C.c_int32), ("identityCheck", C.c_int32),
                ("tracebackBytes", C.c_uint32), ("speculation", C.c_int32), ("overlyPairedKeep", C.c_int32), ("maxPairedBases", C.c_uint64)]


class GappedStats(C.Structure):
    _fields_ = [("anchors", C.c_uint64), ("anchorsExtended", C.c_uint64), ("dpCells", C.c_uint64),
                ("dpRows", C.c_uint64), ("truncated", C.c_uint64), ("speculated", C.c_uint64),
                ("redone", C.c_uint64), ("seconds", C.c_double), ("kernelSeconds", C.c_double * 4),
                ("launches", C.c_uint64), ("dpCellsComputed", C.c_uint64), ("overlyPaired", C.c_uint64)]


assert C.sizeof(Segment) == 48 and C.sizeof(Alignel) == 64

# every symbol include/lastz_b200.h declares
SYMBOLS = ["lzb_backend", "lzb_last_error", "lzb_open", "lzb_close", "lzb_set_scoring",
           "lzb_target_build", "lzb_target_free", "lzb_target_export_index", "lzb_target_limit", "lzb_query_load",
           "lzb_query_free", "lzb_seed_hit_search", "lzb_reduce_to_points", "lzb_gapped_extend",
           "lzb_free_align_list", "lzb_free", "lzb_launch_count"]


def _bind(lib):
    vp = C.c_void_p
    lib.lzb_backend.restype = C.c_char_p
    lib.lzb_last_error.restype = C.c_char_p
    lib.lzb_open.restype = vp
    lib.lzb_open.argtypes = [C.c_int]
    lib.lzb_close.argtypes = [vp]
    lib.lzb_set_scoring.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.c_int32]
    lib.lzb_target_build.restype = vp
    lib.lzb_target_build.argtypes = [vp, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.POINTER(C.c_int8), C.POINTER(Seed), C.c_uint32]
    lib.lzb_target_free.argtypes = [vp]
    lib.lzb_target_export_index.restype = C.c_int64
    lib.lzb_target_export_index.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.lzb_target_limit.argtypes = [vp, C.c_uint32]
    lib.lzb_query_load.restype = vp
    lib.lzb_query_load.argtypes = [vp, C.c_char_p, C.c_uint32]
    lib.lzb_query_free.argtypes = [vp]
    lib.lzb_seed_hit_search.argtypes = [vp, vp, vp, C.POINTER(Seed), C.POINTER(C.c_int8), C.POINTER(SeedParams),
                                        C.POINTER(C.POINTER(Segment)), C.POINTER(C.c_uint64), C.POINTER(SeedStats)]
    lib.lzb_reduce_to_points.argtypes = [vp, vp, vp, C.POINTER(Segment), C.c_uint64]
    lib.lzb_gapped_extend.argtypes = [vp, vp, vp, C.c_char_p, C.c_char_p, C.POINTER(Segment), C.c_uint64,
                                      C.POINTER(GappedParams), C.POINTER(C.POINTER(Alignel)), C.POINTER(GappedStats)]
    lib.lzb_free_align_list.argtypes = [C.POINTER(Alignel)]
    lib.lzb_free.argtypes = [vp]
    lib.lzb_launch_count.restype = C.c_uint64
    lib.lzb_launch_count.argtypes = [vp]
    return lib


def load_product():
    """The CUDA library.  Raises if it has not been built -- there is no other implementation."""
    if not os.path.exists(PRODUCT_LIB):
        raise RuntimeError(f"{PRODUCT_LIB} is missing: run __graft_entry__.build() (nvcc, sm_100a); "
                           "lastz_b200 has no CPU fallback")
    return _bind(C.CDLL(PRODUCT_LIB))


def load_oracle():
    """TEST INFRASTRUCTURE: the CPU restatement (oracle/liblzb_oracle.so)."""
    if not os.path.exists(ORACLE_LIB):
        raise RuntimeError(f"{ORACLE_LIB} is missing: run `make -C oracle`")
    return _bind(C.CDLL(ORACLE_LIB))
