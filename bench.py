#!/usr/bin/env python
"""bench.py -- LASTZ seed-and-extend hot path on B200: seed-hits/s and Y-drop Gcells/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload auto|config3|config4] [--impl ours|reference]

Workloads (BASELINE.json configs, SURVEY.md 8d generator: splitmix64 seed 20260925, ~5 % divergence):
  config3  synthetic 50 Mbp x 50 Mbp, default lastz options, both strands, gapped          -- the 1-GPU line
  config4  synthetic 250 Mbp x 250 Mbp, --chain, gapped, query cut into one interval per GPU -- the multi-GPU line
`--workload auto` (the default) takes config3 on one GPU and config4 on several; the one-GPU line also carries one
pass of config4 (`config4_one_gpu`) so that the multi-GPU numbers have their own single-GPU base.

A "step" is one pass of the hot path over this rank's query interval: for each strand seed_hit_search ->
[reduce_to_chain] -> reduce_to_points -> gapped_extend through the C-ABI.  The target bytes and its position table stay
resident in HBM (built once, outside the timed region, and reported).  The stages run one after the other, so that
the two rates the metric names are those of each stage with the GPU to itself; `other_schedule` in the line times the
same steps with the first strand's gapped stage -- a chain of dependent sweeps that leaves most issue slots idle -- on a
second host thread and context beside the second strand's seed stage (`--overlap` makes that the main line).  After each step the ranks' HSP tables and alignments
are gathered to rank 0 over NCCL.

`value` = raw seed hits of the step / WHOLE step time (both stages of both strands), device-resident query;
`e2e.value` = the same with the query coming from host memory every step and the results read back.  The stage rates
the metric names are in `seed_hits_per_s` (hits / seed-stage time) and `gcells_per_s` (DP cells / gapped-stage time).
Timing: CUDA events either side of the K steps, bracketed by barrier + device sync (max over ranks), the host clock
beside it as a cross-check; per-kernel numbers come from CUDA events recorded by the library on its own stream.
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
# Must be in the environment before torch creates the CUDA context (lzb_open sets it too).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

SEED = 20260925
GAMMA = np.uint64(0x9E3779B97F4A7C15)
WORKLOADS = {
    "config3": dict(L=50_000_000, chain=False,
                    name="BASELINE configs[2]: synthetic 50 Mbp target x ~5%-divergent copy, default lastz options, both strands, gapped"),
    "config4": dict(L=250_000_000, chain=True,
                    name="BASELINE configs[3]: synthetic 250 Mbp target x ~5%-divergent copy, --chain, both strands, gapped, query cut into one interval per GPU"),
}
INT32_LANES_PER_SM = 128          # integer lanes per SM per clock (4 schedulers x 32)


def _splitmix(seed, n, start=0):
    """outputs start+1 .. start+n of splitmix64 started at `seed` (vectorised; wraps modulo 2^64)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + GAMMA * np.arange(start + 1, start + n + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_pair(L, seed=SEED, block=25_000_000):
    """tools/gen_synth.c in numpy, in blocks so that 250 Mbp stays within a few GB: (target bytes, query bytes)."""
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    tparts, qparts = [], []
    for b0 in range(0, L, block):
        n = min(block, L - b0)
        tcode = (_splitmix(seed, n, b0) >> np.uint64(62)).astype(np.uint8)
        r = _splitmix(seed + 1, n, b0)
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
        tparts.append(acgt[tcode].tobytes())
        qparts.append(acgt[q].tobytes())
    return b"".join(tparts), b"".join(qparts)


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
REF_TARGET_BP = 50_000_000        # the reference's sample always runs against (a prefix of) 50 Mbp of target


def cpu_reference(target, query, sample_bp, procs, chain, check_counts=None):
    """Time oracle/_ref/lastz on `procs` query subranges against the first 50 Mbp of the target.

    Stage times by difference: each process's subrange at two lengths (sample_bp and
    2 x sample_bp), with --nogapped for the seed stage and re-run from those HSPs (--segments=) for the gapped stage, so
    that start-up, file loading and the index build cancel.  The metric's numerators (raw seed hits, DP cells) come from
    the reference's own counter build (oracle/_ref/lastz_stats, -Dcollect_stats) on the longer subranges.
    check_counts(ranges) -> (hits, cells), when given, is the PRODUCT on the same subranges: a mismatch is fatal."""
    ref = os.path.join(ROOT, "oracle", "_ref", "lastz")
    counter = os.path.join(ROOT, "oracle", "_ref", "lastz_stats")
    kind = "reference"
    if not os.path.exists(ref):
        ref = counter = os.path.join(ROOT, "oracle", "lastz_oracle")
        kind = "port"
    tlen = min(len(target), REF_TARGET_BP)
    if "fasta" not in _REF_CACHE:                       # the inputs are written once per process
        d = tempfile.mkdtemp(prefix="lzb_bench_")
        _REF_CACHE["fasta"] = (os.path.join(d, "t.fa"), os.path.join(d, "q.fa"))
        write_fasta(_REF_CACHE["fasta"][0], b"t", target[:tlen])
        write_fasta(_REF_CACHE["fasta"][1], b"q", query[:tlen])
    tfa, qfa = _REF_CACHE["fasta"]
    procs = max(1, min(procs, tlen // (2 * sample_bp)))
    stride = tlen // procs                               # spread over the whole sample so that every piece of target is used
    short = [(k * stride + 1, k * stride + sample_bp) for k in range(procs)]
    long_ = [(k * stride + 1, k * stride + 2 * sample_bp) for k in range(procs)]
    opts = ["--chain"] if chain else []

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

    def counts(ranges):
        hits = cells = 0
        ps = [subprocess.Popen([counter, tfa, f"{qfa}[{a}..{b}]", "--stats"] + opts, stdout=subprocess.DEVNULL,
                               stderr=subprocess.PIPE, text=True) for a, b in ranges]   # --stats reports on stderr
        for p in ps:
            for line in p.communicate()[1].splitlines():
                if "raw seed hits:" in line:
                    hits += int(line.split(":")[1].replace(",", ""))
                if "DP cells visited:" in line:
                    cells += int(line.split(":")[1].replace(",", ""))
        return hits, cells

    key = (sample_bp, procs, chain)
    if key not in _REF_CACHE:                            # once per process: warm the file cache, count, cross-check
        run(short, ["--nogapped"] + opts, "short")
        hs, cs = counts(short)
        hl, cl = counts(long_)
        if check_counts is not None and kind == "reference":
            ph, pc = check_counts(long_)
            if (ph, pc) != (hl, cl):
                raise SystemExit(f"FAILURE: bench self-check: the product counts {ph} raw seed hits / {pc} DP cells on the "
                                 f"sample subranges, the reference's counter build {hl} / {cl}")
        _REF_CACHE[key] = ((hs, cs), (hl, cl))
    (hs, cs), (hl, cl) = _REF_CACHE[key]
    t_ns, t_nl = run(short, ["--nogapped"] + opts, "short"), run(long_, ["--nogapped"] + opts, "long")
    t_gs, t_gl = run(short, opts, "short"), run(long_, opts, "long")
    hits, cells = hl - hs, cl - cs
    seed_s, gap_s = t_nl - t_ns, t_gl - t_gs
    note = ""
    if seed_s < 0.05 * t_nl or gap_s < 0.05 * t_gl or cells <= 0:
        # sample too small for differences to rise above process start-up noise: charge whole runs
        # (index build / file loading included), which can only flatter the CPU less
        hits, cells = hl, max(cl, 1)
        seed_s, gap_s = t_nl, max(t_gl, 1e-3)
        note = " [differences below noise: whole-run times used]"
    return {"kind": kind, "cores": procs, "hits": hits, "cells": cells, "index_s": max(t_ns - seed_s, 0.0), "seed_s": seed_s,
            "gapped_s": gap_s, "hits_per_s": hits / seed_s, "cells_per_s": cells / gap_s,
            "self_check": "product hits and DP cells on the sample subranges equal oracle/_ref/lastz_stats" if check_counts and kind == "reference" else None,
            "sample": f"{procs} processes, each query[{sample_bp} bp] and query[{2 * sample_bp} bp] (spread over the sequence) vs the first "
                      f"{tlen} bp of the target, both strands{', --chain' if chain else ''}; stage times are differences between the two lengths "
                      f"(--nogapped for the seed stage; the gapped stage re-run from those HSPs with --segments=, so no index build "
                      f"enters it); numerators from the -Dcollect_stats build" + note}


def whole_job_rate(hits_full, cells_full, hits_per_s, cells_per_s):
    """seed hits per second of whole-job time for a workload with that many hits and cells, at the measured stage rates
    (the linear extrapolation BASELINE.md section 3.5 prescribes)"""
    return hits_full / (hits_full / hits_per_s + cells_full / cells_per_s)


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("LZB_BENCH_WORKLOAD", "auto"), choices=["auto", "config3", "config4"])
    ap.add_argument("--size", type=int, default=0, help="override the workload's sequence length (tests)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--speculation", type=int, default=384, help="anchors in flight per gapped stage (results do not depend on it)")
    ap.add_argument("--cpu-sample", type=int, default=500_000, help="query bp per reference process (x1 and x2)")
    ap.add_argument("--cpu-procs", type=int, default=0, help="reference processes (default: every host core, at most 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--resident-only", action="store_true", help="profiling runs: skip the end-to-end pass (the e2e key then repeats the resident pass and says so)")
    ap.add_argument("--overlap-extend-ctas", type=int, default=2, help="k_extend2 CTAs per SM while it runs beside the other strand's sweeps (0: the library's 4)")
    ap.add_argument("--overlap", action="store_true", help="time the overlapped schedule as the main line (first strand's gapped stage beside the second strand's seed stage)")
    ap.add_argument("--no-overlap", action="store_true", help="(the default since the stage rates are what the metric names) everything one after the other")
    ap.add_argument("--no-config4-base", action="store_true", help="skip the single pass of config4 on the one-GPU line")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wname = args.workload if args.workload != "auto" else ("config3" if world == 1 else "config4")
    wl = dict(WORKLOADS[wname])
    if args.size:
        wl["L"] = args.size
    L = wl["L"]
    ncores = os.cpu_count() or 1
    procs = args.cpu_procs or min(ncores, 64)

    def config_of(w, shards):
        return {"workload": w["name"] if not args.size else f"{w['name']} [length overridden: {w['L']} bp]",
                "generator": "splitmix64 seed 20260925 (SURVEY.md 8d)", "seed": "12of19 + 1 transition",
                "x_drop": 910, "y_drop": 9400, "hsp_threshold": 3000, "traceback_bytes": 80 * 1024 * 1024,
                "chain": bool(w["chain"]), "query_shards": shards,
                "l2": "inputs larger than L2 (index, hit buffers and traceback are GBs)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        target, query = synth_pair(min(L, REF_TARGET_BP))
        # the workload's own numerators, scaled from the sample: random hits grow with L1 x L2, DP cells with L2
        vals = [cpu_reference(target, query, args.cpu_sample, procs, wl["chain"]) for _ in range(max(1, args.steps))]
        hps = float(np.mean([v["hits_per_s"] for v in vals])); cps = float(np.mean([v["cells_per_s"] for v in vals]))
        r = vals[-1]
        tl = min(L, REF_TARGET_BP)
        sample_q = r["cores"] * args.cpu_sample
        hits_full = r["hits"] * (L / tl) * (L / sample_q); cells_full = r["cells"] * (L / sample_q)
        value = whole_job_rate(hits_full, cells_full, hps, cps)
        line = {"impl": "reference", "metric": "seed-hits/s over the whole step (seed + gapped stages); stage rates in seed_hits_per_s and gcells_per_s",
                "value": value, "unit": "hits/s", "seed_hits_per_s": hps, "gcells_per_s": cps / 1e9,
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * hits_full / value, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config_of(wl, args.gpus),
                "extrapolation": f"stage rates measured on the sample; whole-job time = hits/rate + cells/rate with the workload's hits "
                                 f"({hits_full:.3e}) and DP cells ({cells_full:.3e}) scaled from the sample (hits ~ L1 x L2, cells ~ L2; BASELINE.md 3.5); "
                                 f"warm-up: one untimed pass that pages the binary and the inputs in",
                "cpu_baseline": {"value": value, "unit": "hits/s", "seed_hits_per_s": hps, "gcells_per_s": cps / 1e9, "cores": r["cores"], "kind": r["kind"],
                                 "sample": r["sample"], "host_cores_available": ncores},
                "e2e": {"value": value, "unit": "hits/s", "gcells_per_s": cps / 1e9, "seed_hits_per_s": hps, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from lastz_b200 import Engine, default_scoring, parse_seed, reduce_to_chain, revcomp
    from lastz_b200.sharding import gather_to_rank0, pack_alignments, query_interval, to_global

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
    ss = default_scoring()
    seed = parse_seed()
    engA = Engine.product(local)
    engA.set_scoring(ss)
    overlap = args.overlap and not args.no_overlap
    engB = Engine.product(local)                         # a second context (stream, scratch, lanes) for the other strand
    engB.set_scoring(ss)

    lanes_per_scheduler = args.speculation

    def sync():
        if on_gpu:
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def measure(w, steps, warmup, both_passes=True):
        """the timed passes over one workload; returns (resident result, e2e result, static info)"""
        target, query = synth_pair(w["L"])
        lo, hi = query_interval(len(query), rank, world)
        shard = query[lo:hi]
        strands = [(0, shard), (3, revcomp(shard))]
        t0 = time.perf_counter()
        T = engA.build_seed_position_table(target, seed)
        index_s = time.perf_counter() - t0

        def seed_part(eng, sid, s, Q, resident, acc, extend_ctas=0):
            """the seed stage of one strand; returns what the gapped stage needs"""
            w0 = time.perf_counter()
            if not resident:
                Q = eng.load_query(s)
                acc["h2d"] += len(s)
            wl_ = time.perf_counter()
            segs, st = eng.seed_hit_search(T, Q, seed, strand_id=sid, extend_ctas=extend_ctas)
            w1 = time.perf_counter()
            table = segs.copy()
            if w["chain"]:
                segs = reduce_to_chain(segs, ss)
            wc = time.perf_counter()
            acc["seed_wall"] += w1 - w0; acc["chain_wall"] += wc - w1; acc["load_wall"] += wl_ - w0
            acc["hits"] += st.rawSeedHits; acc["seed_s"] += st.seconds; acc["words"] += st.wordsInQuery
            for i in range(12):
                acc["kern"][i] += st.kernelSeconds[i]; acc["kern_n"][i] += st.kernelLaunches[i]
            # the x-drop extension stage: k_extend2 [7] or k_right + k_replay + k_left [8..10], one of each per chunk
            acc["ext_s"] += sum(st.kernelSeconds[i] for i in (7, 8, 9, 10))
            acc["ext_launch"] += st.kernelLaunches[7] + st.kernelLaunches[8]
            acc["bp"] += st.bpExtended; acc["ext"] += st.extensions
            acc["hsps"] += len(table); acc["anchors"] += len(segs)
            return Q, segs, table

        def gapped_part(eng, sid, s, Q, segs, table, resident, acc):
            wc = time.perf_counter()
            anchors = eng.reduce_to_points(T, Q, segs)
            al, gst, _ = eng.gapped_extend(T, Q, target, s, anchors, identity_check=False, speculation=lanes_per_scheduler)
            w2 = time.perf_counter()
            acc["gap_wall"] += w2 - wc
            acc["cells"] += gst.dpCells; acc["cells_computed"] += gst.dpCellsComputed
            acc["gap_s"] += gst.seconds; acc["dp_launches"] += gst.launches; acc["alignments"] += len(al)
            acc["dp_rows"] += gst.dpRows; acc["redone"] += gst.redone; acc["speculated"] += gst.speculated
            acc["h2d"] += 48 * len(segs); acc["d2h"] += 2 * 48 * len(table) + 4 * sum(len(a["ops"]) for a in al)
            if not resident:
                eng.free_query(Q)
            acc["free_wall"] += time.perf_counter() - w2
            acc["tables"].append(to_global(table, lo, hi, len(query), revcomp=(sid != 0)))
            acc["aligns"].append(pack_alignments(al, sid))

        def new_acc():
            return dict(hits=0, cells=0, cells_computed=0, seed_s=0.0, gap_s=0.0, ext_s=0.0, ext_launch=0, bp=0, hsps=0, anchors=0, h2d=0, d2h=0,
                        ext=0, dp_launches=0, seed_wall=0.0, chain_wall=0.0, gap_wall=0.0, load_wall=0.0, free_wall=0.0, gather_wall=0.0, gap_phase_wall=0.0,
                        words=0, kern=[0.0] * 12, kern_n=[0] * 12, alignments=0, dp_rows=0, redone=0, speculated=0, tables=[], aligns=[])

        def step(resident, handles, overlap=overlap):
            accA, accB = new_acc(), new_acc()
            (sidA, sA), (sidB, sB) = strands
            # A strand's gapped stage -- a chain of dependent sweeps that leaves most issue slots idle -- runs on a second host
            # thread and its own context while the main thread is already in the next strand's seed stage.
            QA, segsA, tabA = seed_part(engA, sidA, sA, handles[0] if resident else None, resident, accA)
            if overlap:
                err = []

                def first_gapped():
                    try:
                        gapped_part(engA, sidA, sA, QA, segsA, tabA, resident, accA)
                    except BaseException as e:       # noqa: BLE001  (re-raised on the main thread)
                        err.append(e)
                th = threading.Thread(target=first_gapped)
                th.start()
                try:
                    # beside the first strand's sweeps the extension kernel runs with fewer persistent CTAs per SM, so that the
                    # sweeps (one warp each, latency-bound) keep their issue slots; the seed stage has their whole duration to hide in
                    QB, segsB, tabB = seed_part(engB, sidB, sB, handles[1] if resident else None, resident, accB, extend_ctas=args.overlap_extend_ctas)
                    gapped_part(engB, sidB, sB, QB, segsB, tabB, resident, accB)
                finally:
                    th.join()
                if err:
                    raise err[0]
            else:
                gapped_part(engA, sidA, sA, QA, segsA, tabA, resident, accA)
                QB, segsB, tabB = seed_part(engB, sidB, sB, handles[1] if resident else None, resident, accB)
                gapped_part(engB, sidB, sB, QB, segsB, tabB, resident, accB)
            acc = new_acc()
            for k in acc:
                if k in ("kern", "kern_n"):
                    acc[k] = [a + b for a, b in zip(accA[k], accB[k])]
                else:
                    acc[k] = accA[k] + accB[k]
            acc["first"] = dict(ext_s=accA["ext_s"], ext_launch=accA["ext_launch"], hits=accA["hits"], bp=accA["bp"])
            acc["gap_phase_wall"] = acc["gap_wall"]            # the two strands' gapped calls, each on its own clock
            w3 = time.perf_counter()
            segs_all = gather_to_rank0(np.concatenate(acc["tables"]).view(np.uint8).reshape(-1), dev)
            al_all = gather_to_rank0(np.concatenate(acc["aligns"]).view(np.uint8).reshape(-1), dev)
            acc["gathered_segments"] = sum(len(p) for p in segs_all) // 48 if segs_all is not None else 0
            acc["gathered_alignment_bytes"] = sum(len(p) for p in al_all) if al_all is not None else 0
            acc["gather_wall"] = time.perf_counter() - w3
            acc["tables"], acc["aligns"] = [], []
            return acc

        def timed(resident, steps_, warmup_, ov=overlap):
            handles = [engA.load_query(strands[0][1]), engB.load_query(strands[1][1])] if resident else None
            for _ in range(warmup_):
                step(resident, handles, ov)
            sync()
            l0 = engA.launches() + (engB.launches() if engB is not engA else 0)
            # device clock: CUDA events on torch's stream either side of the steps (every library call of a step blocks the
            # host until its kernels -- on the library's own streams -- are done, so the events bracket all of them); the host
            # clock is kept beside it as a cross-check
            ev0 = ev1 = None
            if on_gpu:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            t0_ = time.perf_counter()
            accs = [step(resident, handles, ov) for _ in range(steps_)]
            if on_gpu:
                ev1.record()
            sync()
            dt = time.perf_counter() - t0_
            if on_gpu:
                dt_dev = ev0.elapsed_time(ev1) / 1e3
                if abs(dt_dev - dt) > 0.02 * dt + 0.005:
                    print(f"[bench] device clock {dt_dev:.4f} s vs host clock {dt:.4f} s", file=sys.stderr)
                dt = dt_dev
            launches = engA.launches() + (engB.launches() if engB is not engA else 0) - l0
            if handles:
                engA.free_query(handles[0]); engB.free_query(handles[1])
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item()), accs, launches

        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        res = timed(True, steps, warmup)
        clocks = sampler.stop() if rank == 0 else None
        e2e = timed(False, steps, 1 if both_passes else 0) if both_passes else None
        # the other schedule, for the record: the first strand's gapped stage on a second thread beside the second strand's seed stage
        other = timed(True, steps, 1, ov=not overlap) if both_passes and w is wl and not args.resident_only else None
        info = dict(target=target, query=query, T=T, index_s=index_s, clocks=clocks, lo=lo, hi=hi, other=other)
        return res, e2e, info

    def total(key, accs_, op=None):
        v = torch.tensor([float(sum(a[key] for a in accs_))], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(v, op=op or dist.ReduceOp.SUM)
        return float(v.item())

    def worst(key, accs_):
        return total(key, accs_, dist.ReduceOp.MAX)

    res_, e2e_, info = measure(wl, args.steps, args.warmup, both_passes=not args.resident_only)
    (dt, accs, launches), (dt_e2e, accs_e2e, _) = res_, (e2e_ if e2e_ is not None else res_)
    target, query, T = info["target"], info["query"], info["T"]
    hits, cells = total("hits", accs), total("cells", accs)
    seed_wall, gap_wall = worst("seed_wall", accs), worst("gap_phase_wall", accs)      # gapped: the strands' gapped calls, each on its own clock
    seed_dev = worst("seed_s", accs)
    total_l = torch.tensor([float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_l, op=dist.ReduceOp.SUM)
    # roofline of the dominant seed-stage kernel (k_extend2), this rank
    ext_s = sum(a["ext_s"] for a in accs); ext_n = max(1, sum(a["ext_launch"] for a in accs))
    my_hits = sum(a["hits"] for a in accs); my_bp = sum(a["bp"] for a in accs)
    bytes_per_hit = 12.0 + 0.5 * my_bp / max(1, my_hits)      # 4 B position + 8 B diagEnd + 0.5 B/column
    peak, peak_src = measured_peak()
    achieved = bytes_per_hit * my_hits / max(ext_s, 1e-12) / 1e9
    # the first strand's launches alone: with the strands overlapped the second strand's k_extend2 runs beside the first
    # strand's Y-drop sweeps (and with fewer CTAs per SM on purpose), which stretches its launches
    f_s = sum(a["first"]["ext_s"] for a in accs); f_n = max(1, sum(a["first"]["ext_launch"] for a in accs))
    f_hits = sum(a["first"]["hits"] for a in accs); f_bp = sum(a["first"]["bp"] for a in accs)
    f_bytes = 12.0 + 0.5 * f_bp / max(1, f_hits)
    f_achieved = f_bytes * f_hits / max(f_s, 1e-12) / 1e9
    # every cross-rank aggregate is computed HERE, on all ranks (collectives must not sit under `if rank == 0`)
    other_line = None
    if info.get("other"):
        dt_o, accs_o, _ = info["other"]
        other_line = {"overlap": not overlap, "ms_per_step": 1e3 * dt_o / args.steps, "value": total("hits", accs_o) / dt_o, "unit": "hits/s",
                      "stage_ms_per_step": {"seed": 1e3 * worst("seed_wall", accs_o) / args.steps, "gapped": 1e3 * worst("gap_phase_wall", accs_o) / args.steps},
                      "what": "the same steps with the first strand's gapped stage on a second host thread beside the second strand's seed stage "
                              "(k_extend2 at %d CTAs per SM meanwhile): the step gets shorter, both stages get slower than alone" % args.overlap_extend_ctas
                              if not overlap else "the same steps with everything one after the other"}
    agg = {"e2e_hits": total("hits", accs_e2e), "e2e_cells": total("cells", accs_e2e),
           "e2e_seed_wall": worst("seed_wall", accs_e2e), "e2e_gap_wall": worst("gap_phase_wall", accs_e2e),
           "hsps": total("hsps", accs), "anchors": total("anchors", accs), "h2d": total("h2d", accs_e2e), "d2h": total("d2h", accs_e2e),
           "cells_computed": total("cells_computed", accs), "alignments": total("alignments", accs),
           "redone": total("redone", accs), "speculated": total("speculated", accs)}

    def per_step(key, accs_):
        return 1e3 * sum(a[key] for a in accs_) / args.steps

    def wall_breakdown(accs_):
        """rank 0's host clock, ms per step, summed over the two strands (which overlap in time when --no-overlap is not given)"""
        return {"load_query": per_step("load_wall", accs_), "seed_call": per_step("seed_wall", accs_) - per_step("load_wall", accs_),
                "seed_device": per_step("seed_s", accs_), "chain": per_step("chain_wall", accs_), "peaks_and_gapped_call": per_step("gap_wall", accs_),
                "free_query": per_step("free_wall", accs_), "gather_to_rank0": per_step("gather_wall", accs_)}

    KNAMES = ["k_query_words", "k_count_hits", "k_slot_count", "cub scan", "k_expand", "cub radix sort", "k_bucket_bounds",
              "k_extend2", "k_right", "k_replay", "k_left", "bucket order (k_bucket_sizes + cub sort)"]
    kern = [sum(a["kern"][i] for a in accs) for i in range(12)]
    kern_n = [sum(a["kern_n"][i] for a in accs) for i in range(12)]
    my_words = sum(a["words"] for a in accs)
    traffic_per_hit = None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp):
            traffic_per_hit = json.load(open(tp)).get("k_extend_dram_bytes_per_hit")
            traffic_src = name
            break

    base4 = None
    if world == 1 and wname == "config3" and not args.no_config4_base and not args.size:
        # the multi-GPU line's own single-GPU base: one timed pass of config4 (one warm-up pass before it)
        engA.free_position_table(T)
        T = None
        (dt4, accs4, _), _, info4 = measure(WORKLOADS["config4"], 1, 1, both_passes=False)
        h4 = sum(a["hits"] for a in accs4); c4 = sum(a["cells"] for a in accs4)
        base4 = {"workload": WORKLOADS["config4"]["name"], "value": h4 / dt4, "unit": "hits/s", "ms_per_step": 1e3 * dt4, "steps": 1, "warmup": 1,
                 "seed_hits_per_s": h4 / max(sum(a["seed_wall"] for a in accs4), 1e-12),
                 "gcells_per_s": c4 / max(sum(a["gap_phase_wall"] for a in accs4), 1e-12) / 1e9,
                 "counts_per_step": {"raw_seed_hits": h4, "dp_cells": c4, "hsps": sum(a["hsps"] for a in accs4),
                                     "anchors_after_chain": sum(a["anchors"] for a in accs4), "alignments": sum(a["alignments"] for a in accs4)},
                 "index_build_once_ms": 1e3 * info4["index_s"]}
        engA.free_position_table(info4["T"])
        target, query = synth_pair(L)                    # back to this line's workload for the CPU leg
        T = engA.build_seed_position_table(target, seed)

    if rank == 0:
        sm_mhz = (info["clocks"] or {}).get("sm_mhz") or 1965
        line = {"metric": "seed-hits/s over the whole step (seed + gapped stages); stage rates in seed_hits_per_s and gcells_per_s",
                "value": hits / dt, "unit": "hits/s",
                "seed_hits_per_s": hits / seed_wall, "gcells_per_s": cells / gap_wall / 1e9,
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "int32", "data": "synthetic", "config": config_of(wl, world),
                "stage_ms_per_step": {"seed": 1e3 * seed_wall / args.steps, "seed_device_events": 1e3 * seed_dev / args.steps,
                                      "gapped": 1e3 * gap_wall / args.steps, "index_build_once": 1e3 * info["index_s"],
                                      "note": "wall time of the blocking calls, summed over the two strands, max over ranks" +
                                              ("; the first strand's gapped call overlaps the second strand's seed call, so the stages add up to "
                                               "more than ms_per_step and both are slower than alone" if overlap else "")},
                "counts_per_step": {"raw_seed_hits": hits / args.steps, "dp_cells": cells / args.steps,
                                    "dp_cells_incl_discarded_speculation": agg["cells_computed"] / args.steps,
                                    "hsps": agg["hsps"] / args.steps, "anchors": agg["anchors"] / args.steps, "alignments": agg["alignments"] / args.steps,
                                    "sweeps_started_speculatively": agg["speculated"] / args.steps, "sweeps_resumed_or_redone": agg["redone"] / args.steps,
                                    "segments_gathered": accs[-1]["gathered_segments"], "alignment_bytes_gathered": accs[-1]["gathered_alignment_bytes"]},
                "timing": "CUDA events either side of the K steps (barrier + synchronize both sides), max over ranks, host clock as cross-check; "
                          "kernels timed by CUDA events on the library's stream; strands overlapped: " + str(overlap),
                "clocks": info["clocks"], "gpu_launches": int(total_l.item()),
                "e2e": {"value": agg["e2e_hits"] / dt_e2e, "unit": "hits/s",
                        "seed_hits_per_s": agg["e2e_hits"] / agg["e2e_seed_wall"], "gcells_per_s": agg["e2e_cells"] / agg["e2e_gap_wall"] / 1e9,
                        "ms_per_step": 1e3 * dt_e2e / args.steps,
                        "what": "NOT MEASURED (--resident-only): a copy of the device-resident pass" if args.resident_only else
                                "the same step with the query strands copied from host memory inside the timed region (lzb_query_load) and HSP tables + "
                                "edit scripts read back",
                        "h2d_bytes_per_step": int(agg["h2d"] / args.steps),
                        "d2h_bytes_per_step": int(agg["d2h"] / args.steps)},
                "roofline": {"kernel": "k_extend2 (bucket replay + x-drop extension, DESIGN.md K3): the seed-hit kernel north_star's HBM target names",
                             "bound": "hbm", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None if traffic_per_hit is None else traffic_per_hit * my_hits / ext_n,
                             "traffic_source": None if traffic_per_hit is None else
                             f"dram__bytes_read+write per hit from the ncu --set full capture in profiles/ ({traffic_src}) x hits per launch here",
                             "peak_source": peak_src,
                             "bytes_per_hit": bytes_per_hit, "launches": ext_n, "avg_launch_ms": 1e3 * ext_s / ext_n,
                             "first_strand_alone": {"achieved": f_achieved, "frac": f_achieved / peak, "launches": f_n, "avg_launch_ms": 1e3 * f_s / f_n,
                                                    "note": "the launches of the first strand, which run with the GPU to themselves; achieved/frac above "
                                                            "average over all launches of the timed region, the second strand's beside the first "
                                                            "strand's Y-drop sweeps included"} if overlap else None},
                "wall_ms_per_step": {"resident": wall_breakdown(accs), "e2e": wall_breakdown(accs_e2e)},
                "other_schedule": other_line}
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
        my_cells = sum(a["cells_computed"] for a in accs); gap_my = sum(a["gap_phase_wall"] for a in accs)
        int_peak = 148 * INT32_LANES_PER_SM * sm_mhz * 1e6
        rk.append({"kernel": "k_ydrop_warp / k_ydrop_mw (Y-drop DP + traceback; one CTA per one-sided sweep, hundreds of sweeps in flight)",
                   "launches": sum(a["dp_launches"] for a in accs), "gapped_stage_ms_per_step": 1e3 * gap_my / args.steps,
                   "bound": "dependent integer issue per DP row (max-plus recurrence + two block-wide scans), not HBM",
                   "achieved_gbs": 1.0 * my_cells / max(gap_my, 1e-12) / 1e9, "frac_of_hbm_peak": 1.0 * my_cells / max(gap_my, 1e-12) / 1e9 / peak,
                   "int32_frac": 12.0 * my_cells / max(gap_my, 1e-12) / int_peak,
                   "int32_note": f"12 integer ops per cell (BASELINE.md 4) x cells computed / gapped-stage time, against 148 SMs x {INT32_LANES_PER_SM} lanes x "
                                 f"{sm_mhz} MHz = {int_peak / 1e12:.1f} Tops/s",
                   "note": "1 traceback byte per cell is the only algorithmic HBM traffic",
                   "cells_computed_per_step": my_cells / args.steps, "rows_per_step": sum(a["dp_rows"] for a in accs) / args.steps})
        line["roofline_kernels"] = rk
        if base4 is not None:
            line["config4_one_gpu"] = base4
        if world == 1 and not args.no_cpu_baseline:
            def product_counts(ranges):
                h = c = 0
                tl = min(len(target), REF_TARGET_BP)
                Ts = T if tl == len(target) else engA.build_seed_position_table(target[:tl], seed)
                for a, b in ranges:
                    sub = query[a - 1:b]
                    for sid, s in ((0, sub), (3, revcomp(sub))):
                        Q = engA.load_query(s)
                        segs, st = engA.seed_hit_search(Ts, Q, seed, strand_id=sid)
                        if wl["chain"]:
                            segs = reduce_to_chain(segs, ss)
                        anchors = engA.reduce_to_points(Ts, Q, segs)
                        _, gst, _ = engA.gapped_extend(Ts, Q, target[:tl], s, anchors, speculation=args.speculation)
                        h += st.rawSeedHits; c += gst.dpCells
                        engA.free_query(Q)
                if Ts is not T:
                    engA.free_position_table(Ts)
                return h, c
            r = cpu_reference(target, query, args.cpu_sample, procs, wl["chain"], product_counts)
            cpu_value = whole_job_rate(hits / args.steps, cells / args.steps, r["hits_per_s"], r["cells_per_s"])
            line["cpu_baseline"] = {"value": cpu_value, "unit": "hits/s", "seed_hits_per_s": r["hits_per_s"], "gcells_per_s": r["cells_per_s"] / 1e9,
                                    "cores": r["cores"], "kind": r["kind"], "sample": r["sample"], "self_check": r["self_check"],
                                    "host_cores_available": ncores,
                                    "whole_job": "stage rates measured on the sample, whole-job time = this step's hits / rate + this step's DP cells / rate"}
            line["speedup_vs_cpu_baseline"] = {"whole_step_e2e": line["e2e"]["value"] / cpu_value,
                                               "seed_stage_e2e": line["e2e"]["seed_hits_per_s"] / r["hits_per_s"],
                                               "gapped_stage_e2e": line["e2e"]["gcells_per_s"] / (r["cells_per_s"] / 1e9),
                                               "note": f"against {r['cores']} host cores running the unmodified reference; reported, not a target"}
        print(json.dumps(line))
    if T is not None:
        engA.free_position_table(T)
    if engB is not engA:
        engB.close()
    engA.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    # stdout carries exactly one line, the JSON: anything a library prints there meanwhile (NCCL's version banner, ...)
    # goes to stderr
    _out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    _print = print

    def print(*a, **k):                    # noqa: A001  (module-level shadow on purpose: main() prints only the JSON line to stdout)
        if k.get("file") is None:
            k["file"] = _out
            k["flush"] = True
        _print(*a, **k)
    sys.exit(main())
