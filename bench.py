#!/usr/bin/env python
"""bench.py -- LASTZ seed-and-extend hot path on B200: seed-hits/s and Y-drop Gcells/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size L] [--impl ours|reference]

Workload (BASELINE.json configs[2]): synthetic L = 50 Mbp target x its ~5 %-divergent copy (the
SURVEY.md 8d splitmix64 generator, seed 20260925), default options (12of19 + 1 transition,
x-drop 910, K = L = 3000, entropy, Y-drop 9400, 80 MiB traceback), both strands.
A "step" is one pass of the hot path over this rank's query interval: for each strand,
seed_hit_search -> reduce_to_points -> gapped_extend through the C-ABI.  The target bytes and its
position table stay resident in HBM (built once, outside the timed region, and reported).
With N ranks the QUERY is cut into N equal intervals (strong scaling, reference-visible cuts
`q.fa[a..b]`); after each step the ranks' HSP segment tables are gathered with one NCCL
all_gather over NVLink.

`value` = raw seed hits per second of seed-stage time (the metric's first half); `gcells_per_s` =
DP cells per second of gapped-stage time (its second half).  `ms_per_step` covers the whole step.
Timing is the host clock around blocking C-ABI calls bracketed by barrier + device sync (max over
ranks); per-kernel numbers come from CUDA events recorded by the library on its own stream.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the gapped stage runs one stream per speculation lane; give each its own hardware queue.
# Must be in the environment before torch creates the CUDA context (lzb_open sets it too).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

SEED = 20260925
GAMMA = np.uint64(0x9E3779B97F4A7C15)


def _splitmix(seed, n):
    """n outputs of splitmix64 started at `seed` (vectorised; wraps modulo 2^64)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + GAMMA * np.arange(1, n + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_pair(L, seed=SEED):
    """tools/gen_synth.c in numpy: (target bytes, query bytes)."""
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    tcode = (_splitmix(seed, L) >> np.uint64(62)).astype(np.uint8)
    r = _splitmix(seed + 1, L)
    u = (r & np.uint64(0xFFFFFF)).astype(np.int64)
    hi = (r >> np.uint64(24))
    sub = u < 671089
    dele = (u >= 671089) & (u < 754975)
    ins = (u >= 754975) & (u < 838861)
    first = np.where(sub, (tcode + 1 + (hi % np.uint64(3)).astype(np.uint8)) & 3, tcode).astype(np.uint8)
    cnt = np.where(dele, 0, np.where(ins, 2, 1)).astype(np.int64)
    off = np.cumsum(cnt) - cnt
    q = np.empty(int(cnt.sum()), dtype=np.uint8)
    keep = ~dele
    q[off[keep]] = first[keep]
    q[off[ins] + 1] = (hi[ins] & np.uint64(3)).astype(np.uint8)
    return acgt[tcode].tobytes(), acgt[q].tobytes()


def write_fasta(path, name, seq):
    with open(path, "wb") as f:
        f.write(b">" + name + b"\n")
        f.write(b"\n".join(seq[i:i + 60] for i in range(0, len(seq), 60)))
        f.write(b"\n")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------
# the reference / CPU baseline arm: unmodified lastz (oracle/_ref) on the host cores
# --------------------------------------------------------------------------------------------
_REF_CACHE = {}


def cpu_reference(target, query, sample_bp, procs, hits_cells_fn):
    """Time oracle/_ref/lastz on `procs` query subranges of `sample_bp` each against the full target.

    Stage times by difference: the same query prefix at two lengths, with and
    without --nogapped, so index build and start-up cancel.  hits/cells for the sample come from
    hits_cells_fn (the product, proven equal by the parity tests) or from the counter build."""
    ref = os.path.join(ROOT, "oracle", "_ref", "lastz")
    kind = "reference"
    if not os.path.exists(ref):
        ref = os.path.join(ROOT, "oracle", "lastz_oracle")
        kind = "port"
    if "fasta" not in _REF_CACHE:                       # the inputs are written once per process
        d = tempfile.mkdtemp(prefix="lzb_bench_")
        _REF_CACHE["fasta"] = (os.path.join(d, "t.fa"), os.path.join(d, "q.fa"))
        write_fasta(_REF_CACHE["fasta"][0], b"t", target)
        write_fasta(_REF_CACHE["fasta"][1], b"q", query)
    tfa, qfa = _REF_CACHE["fasta"]
    procs = max(1, min(procs, len(query) // (2 * sample_bp)))
    short = [(k * 2 * sample_bp + 1, k * 2 * sample_bp + sample_bp) for k in range(procs)]
    long_ = [(k * 2 * sample_bp + 1, (k + 1) * 2 * sample_bp) for k in range(procs)]

    def run(ranges, extra, tag=None):
        """one reference process per range, concurrently; tag: HSPs go to / anchors come from a segments file per range"""
        cmds = []
        for k, (a, b) in enumerate(ranges):
            cmd = [ref, tfa, f"{qfa}[{a}..{b}]"] + extra
            if tag and "--nogapped" in extra:
                cmd += ["--format=segments", f"--output={tfa}.{tag}.{k}.segments"]
            elif tag:
                cmd += [f"--segments={tfa}.{tag}.{k}.segments"]
            cmds.append(cmd)
        t0 = time.perf_counter()
        ps = [subprocess.Popen(c, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for c in cmds]
        for p in ps:
            p.wait()
        return time.perf_counter() - t0

    # Seed stage: --nogapped runs (their HSPs are kept as segments files); gapped stage: the reference re-run from
    # those files with --segments=, which skips index build and seed search (src/Makefile:384 base_test_segments),
    # so the 6 s index build of a 50 Mbp target and its run-to-run noise never enter the gapped-stage time.
    # Start-up, file loading and index build cancel in the differences between the two query lengths.  All four
    # runs are timed afresh in every step (a cached first measurement would carry the cold file cache into every
    # later step); only the hit/cell counts of the sample are taken once per process.
    key = (sample_bp, procs)
    if key not in _REF_CACHE:
        run(short, ["--nogapped"], "short")             # untimed: pages the binary and the two FASTA files in
        _REF_CACHE[key] = (hits_cells_fn(short), hits_cells_fn(long_))
    (hs, cs), (hl, cl) = _REF_CACHE[key]
    t_ns, t_nl = run(short, ["--nogapped"], "short"), run(long_, ["--nogapped"], "long")
    t_gs, t_gl = run(short, [], "short"), run(long_, [], "long")
    hits, cells = hl - hs, cl - cs
    seed_s = t_nl - t_ns
    gap_s = t_gl - t_gs
    note = ""
    if seed_s < 0.05 * t_nl or gap_s < 0.05 * t_gl:
        # sample too small for differences to rise above process start-up noise: charge whole runs
        # (index build / file loading included), which can only flatter the CPU less
        hits, cells = hl, cl
        seed_s, gap_s = t_nl, max(t_gl, 1e-3)
        note = " [differences below noise: whole-run times used]"
    return {"kind": kind, "cores": procs, "hits": hits, "cells": cells, "index_s": t_ns - seed_s, "seed_s": seed_s,
            "gapped_s": gap_s, "hits_per_s": hits / seed_s, "gcells_per_s": cells / gap_s / 1e9,
            "sample": f"{procs} processes, each query[{sample_bp} bp] and query[{2 * sample_bp} bp] vs the full "
                      f"{len(target)} bp target, both strands; stage times are differences between the two lengths "
                      f"(--nogapped for the seed stage; the gapped stage re-run from those HSPs with --segments=, "
                      f"so no index build enters it)" + note}


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=50_000_000)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--speculation", type=int, default=256)
    ap.add_argument("--cpu-sample", type=int, default=125_000, help="query bp per reference process (x1 and x2)")
    ap.add_argument("--cpu-procs", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    L = args.size
    config = {"workload": f"synthetic {L} bp target x ~5%-divergent copy, default lastz options, both strands, gapped",
              "generator": "splitmix64 seed 20260925 (SURVEY.md 8d)", "seed": "12of19 + 1 transition",
              "x_drop": 910, "y_drop": 9400, "hsp_threshold": 3000, "traceback_bytes": 80 * 1024 * 1024,
              "query_shards": world, "l2": "inputs larger than L2 (index + hit buffers are GBs)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        target, query = synth_pair(L)
        ncores = os.cpu_count() or 1
        procs = min(ncores, 64)
        counter = os.path.join(ROOT, "oracle", "_ref", "lastz_stats")

        def count_with_stats(ranges):
            # the counter build prints the metric numerators itself
            tfa, qfa = _REF_CACHE["fasta"]               # written by cpu_reference before it asks for counts
            hits = cells = 0
            ps = [subprocess.Popen([counter, tfa, f"{qfa}[{a}..{b}]", "--stats"], stdout=subprocess.DEVNULL,
                                   stderr=subprocess.PIPE, text=True) for a, b in ranges]   # --stats reports on stderr
            for p in ps:
                out = p.communicate()[1]
                for line in out.splitlines():
                    if "raw seed hits:" in line:
                        hits += int(line.split(":")[1].replace(",", ""))
                    if "DP cells visited:" in line:
                        cells += int(line.split(":")[1].replace(",", ""))
            return hits, cells
        vals = []
        for _ in range(args.warmup + args.steps):
            r = cpu_reference(target, query, args.cpu_sample, procs, count_with_stats)
            vals.append(r)
        r = vals[-1]
        hps = float(np.mean([v["hits_per_s"] for v in vals[args.warmup:]]))
        line = {"impl": "reference", "metric": "seed-hits/s (seed stage); Gcells/s in gcells_per_s", "value": hps,
                "unit": "hits/s", "gcells_per_s": float(np.mean([v["gcells_per_s"] for v in vals[args.warmup:]])),
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * (r["index_s"] + r["seed_s"] + r["gapped_s"]), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": hps, "unit": "hits/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": hps, "unit": "hits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from lastz_b200 import Engine, default_scoring, parse_seed, revcomp

    # LZB_BENCH_DEVICE=cpu exists for tests/test_bench_dryrun.py only: it runs this very control flow (sharding,
    # collectives, aggregation, the JSON line) with world_size 2 over gloo, the test substituting its own engine.
    # The product engine needs a GPU; nothing here falls back to anything.
    on_gpu = os.environ.get("LZB_BENCH_DEVICE", "cuda") != "cpu"
    dev = torch.device("cuda", local) if on_gpu else torch.device("cpu")
    if world > 1:
        if on_gpu:
            dist.init_process_group("nccl", device_id=dev)
        else:
            dist.init_process_group("gloo")
    if on_gpu:
        torch.cuda.set_device(local)
    target, query = synth_pair(L)
    lo, hi = rank * len(query) // world, (rank + 1) * len(query) // world
    shard = query[lo:hi]
    strands = [(0, shard), (3, revcomp(shard))]

    eng = Engine.product(local)
    eng.set_scoring(default_scoring())
    seed = parse_seed()
    t0 = time.perf_counter()
    T = eng.build_seed_position_table(target, seed)
    index_s = time.perf_counter() - t0

    def sync():
        if on_gpu:
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def gather_segments(tables):
        """the one exchange step: every rank's HSP table to all ranks over NCCL"""
        from lastz_b200.sharding import gather_segment_tables
        parts = gather_segment_tables(np.concatenate(tables), dev)
        return sum(len(p) for p in parts)

    def step(resident, handles=None):
        acc = dict(hits=0, cells=0, cells_computed=0, seed_s=0.0, gap_s=0.0, ext_s=0.0, ext_launch=0, bp=0, hsps=0, h2d=0, d2h=0,
                   ext=0, dp_kernel_s=0.0, dp_launches=0, seed_wall=0.0, gap_wall=0.0, load_wall=0.0, free_wall=0.0, gather_wall=0.0,
                   words=0, kern=[0.0] * 12, kern_n=[0] * 12)
        tables = []
        for k, (sid, s) in enumerate(strands):
            w0 = time.perf_counter()
            Q = handles[k] if resident else eng.load_query(s)
            wl = time.perf_counter()
            if not resident:
                acc["h2d"] += len(s)
            segs, st = eng.seed_hit_search(T, Q, seed, strand_id=sid)
            w1 = time.perf_counter()
            tables.append(segs.copy())
            anchors = eng.reduce_to_points(T, Q, segs)
            al, gst, _ = eng.gapped_extend(T, Q, target, s, anchors, identity_check=False, speculation=args.speculation)
            w2 = time.perf_counter()
            acc["seed_wall"] += w1 - w0; acc["gap_wall"] += w2 - w1; acc["load_wall"] += wl - w0
            acc["hits"] += st.rawSeedHits; acc["seed_s"] += st.seconds; acc["words"] += st.wordsInQuery
            for i in range(12):
                acc["kern"][i] += st.kernelSeconds[i]; acc["kern_n"][i] += st.kernelLaunches[i]
            # the x-drop extension stage: k_extend2 / k_extend [7] or k_right + k_replay + k_left [8..10], one of each per chunk
            acc["ext_s"] += sum(st.kernelSeconds[i] for i in (7, 8, 9, 10))
            acc["ext_launch"] += st.kernelLaunches[7] + st.kernelLaunches[8]
            acc["bp"] += st.bpExtended; acc["ext"] += st.extensions
            acc["hsps"] += len(segs); acc["cells"] += gst.dpCells; acc["cells_computed"] += gst.dpCellsComputed
            acc["gap_s"] += gst.seconds
            acc["dp_kernel_s"] += gst.kernelSeconds[0]; acc["dp_launches"] += gst.launches
            acc["h2d"] += 48 * len(segs); acc["d2h"] += 2 * 48 * len(segs) + 4 * sum(len(a["ops"]) for a in al)
            if not resident:
                eng.free_query(Q)
            acc["free_wall"] += time.perf_counter() - w2
        w3 = time.perf_counter()
        acc["gathered"] = gather_segments(tables)
        acc["gather_wall"] = time.perf_counter() - w3
        return acc

    def timed(resident):
        handles = [eng.load_query(s) for _, s in strands] if resident else None
        for _ in range(args.warmup):
            step(resident, handles)
        sync()
        l0 = eng.launches()
        t0 = time.perf_counter()
        accs = [step(resident, handles) for _ in range(args.steps)]
        sync()
        dt = time.perf_counter() - t0
        launches = eng.launches() - l0
        if handles:
            for h in handles:
                eng.free_query(h)
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()), accs, launches

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dt, accs, launches = timed(True)
    clocks = sampler.stop() if rank == 0 else None
    dt_e2e, accs_e2e, _ = timed(False)

    def total(key, accs_):
        v = torch.tensor([float(sum(a[key] for a in accs_))], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
        return float(v.item())

    def worst(key, accs_):
        v = torch.tensor([float(sum(a[key] for a in accs_))], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    hits, cells = total("hits", accs), total("cells", accs)
    seed_s, gap_s = worst("seed_s", accs), worst("gap_s", accs)
    total_l = torch.tensor([float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_l, op=dist.ReduceOp.SUM)
    # roofline of the dominant seed-stage kernel (k_extend), this rank
    ext_s = sum(a["ext_s"] for a in accs); ext_n = max(1, sum(a["ext_launch"] for a in accs))
    my_hits = sum(a["hits"] for a in accs); my_bp = sum(a["bp"] for a in accs)
    bytes_per_hit = 12.0 + 0.5 * my_bp / max(1, my_hits)      # 4 B position + 8 B diagEnd + 0.5 B/column
    peak, peak_src = measured_peak()
    achieved = bytes_per_hit * my_hits / max(ext_s, 1e-12) / 1e9
    # every cross-rank aggregate is computed HERE, on all ranks (collectives must not sit under `if rank == 0`)
    agg = {"e2e_hits": total("hits", accs_e2e), "e2e_cells": total("cells", accs_e2e),
           "e2e_seed_wall": worst("seed_wall", accs_e2e), "e2e_gap_wall": worst("gap_wall", accs_e2e),
           "hsps": total("hsps", accs), "h2d": total("h2d", accs_e2e), "d2h": total("d2h", accs_e2e),
           "cells_computed": total("cells_computed", accs)}
    e2e_hits = agg["e2e_hits"]

    def per_step(key, accs_):
        return 1e3 * sum(a[key] for a in accs_) / args.steps

    def wall_breakdown(accs_):
        """rank 0's host clock, ms per step: where the time outside the kernels goes"""
        return {"load_query": per_step("load_wall", accs_), "seed_call": per_step("seed_wall", accs_) - per_step("load_wall", accs_),
                "seed_device": per_step("seed_s", accs_), "peaks_and_gapped_call": per_step("gap_wall", accs_),
                "free_query": per_step("free_wall", accs_), "gather_segments": per_step("gather_wall", accs_)}

    # per-kernel view of the seed stage + the DP kernel, this rank (CUDA events on the library's streams)
    KNAMES = ["k_query_words", "k_count_hits", "k_slot_count", "cub scan", "k_expand", "cub radix sort", "k_bucket_bounds",
              "k_extend2", "k_right", "k_replay", "k_left", "bucket order (k_bucket_sizes + cub sort)"]
    kern = [sum(a["kern"][i] for a in accs) for i in range(12)]
    kern_n = [sum(a["kern_n"][i] for a in accs) for i in range(12)]
    my_words = sum(a["words"] for a in accs)
    traffic_per_hit = None
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp):
        traffic_per_hit = json.load(open(tp)).get("k_extend_dram_bytes_per_hit")

    if rank == 0:
        line = {"metric": "seed-hits/s (seed stage); Gcells/s in gcells_per_s", "value": hits / seed_s, "unit": "hits/s",
                "gcells_per_s": cells / gap_s / 1e9, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "int32", "data": "synthetic", "config": config,
                "stage_ms_per_step": {"seed": 1e3 * seed_s / args.steps, "gapped": 1e3 * gap_s / args.steps,
                                      "index_build_once": 1e3 * index_s},
                "counts_per_step": {"raw_seed_hits": hits / args.steps, "dp_cells": cells / args.steps,
                                    "dp_cells_incl_discarded_speculation": agg["cells_computed"] / args.steps,
                                    "hsps": agg["hsps"] / args.steps, "segments_gathered": accs[-1]["gathered"]},
                "timing": "host clock around blocking C-ABI calls, barrier+sync both sides, max over ranks; "
                          "kernels timed by CUDA events on the library's stream",
                "clocks": clocks, "gpu_launches": int(total_l.item()),
                "e2e": {"value": e2e_hits / agg["e2e_seed_wall"],
                        "unit": "hits/s (host clock: H2D copy of the query from host memory + lzb_seed_hit_search incl. D2H of the HSP table)",
                        "gcells_per_s": agg["e2e_cells"] / agg["e2e_gap_wall"] / 1e9,
                        "ms_per_step": 1e3 * dt_e2e / args.steps,
                        "h2d_bytes_per_step": int(agg["h2d"] / args.steps),
                        "d2h_bytes_per_step": int(agg["d2h"] / args.steps)},
                "roofline": {"kernel": "k_extend2 (bucket replay + x-drop extension, DESIGN.md K3): the dominant kernel of the seed stage, "
                                       "i.e. of the time `value` is measured on; the step as a whole is dominated by k_ydrop_mw, see roofline_kernels",
                             "bound": "hbm", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None if traffic_per_hit is None else traffic_per_hit * my_hits / ext_n,
                             "traffic_source": None if traffic_per_hit is None else
                             "dram__bytes_read+write per hit from the ncu --set full capture in profiles/ (r01_traffic.json) x hits per launch here",
                             "peak_source": peak_src,
                             "bytes_per_hit": bytes_per_hit, "launches": ext_n, "avg_launch_ms": 1e3 * ext_s / ext_n},
                "wall_ms_per_step": {"resident": wall_breakdown(accs), "e2e": wall_breakdown(accs_e2e)}}
        # the other kernels against the same HBM peak (algorithmic bytes as in DESIGN.md section 4)
        V = seed.numFlips + 1 if seed.withTrans == 1 else 1
        alg = {4: 4.0 * V * my_words + 16.0 * my_hits,            # k_expand: index probes + 4 B position read + 12 B record written
               5: 2 * 2 * 12.0 * my_hits,                         # radix sort: 2 passes x (read + write) x 12 B (16 hash bits)
               2: 4.0 * V * my_words + 4.0 * V * my_words}        # k_slot_count: probes + one count per slot
        rk = []
        for i in range(12):
            if kern_n[i] == 0:
                continue
            ent = {"kernel": KNAMES[i], "launches": kern_n[i], "ms_per_step": 1e3 * kern[i] / args.steps,
                   "share_of_seed_stage": kern[i] / max(sum(kern), 1e-12)}
            if i in alg:
                ent["achieved_gbs"] = alg[i] / max(kern[i], 1e-12) / 1e9
                ent["frac_of_hbm_peak"] = ent["achieved_gbs"] / peak
            rk.append(ent)
        dp_s = sum(a["dp_kernel_s"] for a in accs); my_cells = sum(a["cells_computed"] for a in accs)
        rk.append({"kernel": "k_ydrop_mw (Y-drop DP + traceback; one launch = the two one-sided DPs of an anchor, up to "
                             f"{args.speculation} launches in flight)", "launches": sum(a["dp_launches"] for a in accs),
                   "lane_ms_per_step": 1e3 * dp_s / args.steps, "bound": "latency (integer max-plus recurrence per row), not HBM",
                   "achieved_gbs": 1.0 * my_cells / max(dp_s, 1e-12) / 1e9, "frac_of_hbm_peak": 1.0 * my_cells / max(dp_s, 1e-12) / 1e9 / peak,
                   "note": "1 traceback byte per cell is the only algorithmic HBM traffic; lane time is summed over concurrent launches",
                   "cells_computed_per_step": my_cells / args.steps})
        line["roofline_kernels"] = rk
        if world == 1 and not args.no_cpu_baseline:
            def hc(ranges):
                h = c = 0
                for a, b in ranges:
                    sub = query[a - 1:b]
                    for sid, s in ((0, sub), (3, revcomp(sub))):
                        Q = eng.load_query(s)
                        segs, st = eng.seed_hit_search(T, Q, seed, strand_id=sid)
                        anchors = eng.reduce_to_points(T, Q, segs)
                        _, gst, _ = eng.gapped_extend(T, Q, target, s, anchors, speculation=args.speculation)
                        h += st.rawSeedHits; c += gst.dpCells
                        eng.free_query(Q)
                return h, c
            ncores = os.cpu_count() or 1
            r = cpu_reference(target, query, args.cpu_sample, min(args.cpu_procs, ncores), hc)
            line["cpu_baseline"] = {"value": r["hits_per_s"], "unit": "hits/s", "gcells_per_s": r["gcells_per_s"],
                                    "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                    "host_cores_available": ncores}
        print(json.dumps(line))
    eng.free_position_table(T)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
