"""bench.py's control flow without a GPU: the real main() -- query sharding, the all_gather of the segment
tables, max/sum aggregation over ranks, the single JSON line from rank 0 -- run with world_size 2 over gloo
(LZB_BENCH_DEVICE=cpu), the TEST substituting the oracle engine for the product engine (tests may use the
oracle; bench.py itself never does outside its cpu_baseline leg).  Catches what a GPU-less container
otherwise cannot: a collective that only some ranks reach, a rank-0-only aggregate, a malformed line."""
import json
import os
import subprocess
import sys

from conftest import ROOT

WORKER = r"""
import os, sys, time
sys.path.insert(0, %(root)r)
os.environ["LZB_BENCH_DEVICE"] = "cpu"
import lastz_b200
from lastz_b200 import Engine

class TimedOracle(Engine):
    # the oracle reports no device times; fill them with the host clock so that rates are finite
    def seed_hit_search(self, *a, **k):
        t0 = time.perf_counter(); segs, st = super().seed_hit_search(*a, **k); dt = time.perf_counter() - t0
        st.seconds = dt; st.kernelSeconds[7] = dt; st.kernelLaunches[7] = 1
        return segs, st
    def gapped_extend(self, *a, **k):
        t0 = time.perf_counter(); al, gst, x = super().gapped_extend(*a, **k); gst.seconds = time.perf_counter() - t0
        return al, gst, x

Engine.product = classmethod(lambda cls, device=0: TimedOracle(lastz_b200.capi.load_oracle(), 0))
import bench
sys.argv = ["bench.py", "--gpus", %(world)r, "--steps", "1", "--warmup", "1", "--size", "150000", "--no-cpu-baseline"]
sys.exit(bench.main())
"""


def _run(tmp_path, world, port):
    script = tmp_path / f"worker{world}.py"
    script.write_text(WORKER % dict(root=ROOT, world=str(world)))
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-3000:]
    lines = [[l for l in o.splitlines() if l.startswith("{")] for o, _ in outs]
    assert len(lines[0]) == 1 and all(len(l) == 0 for l in lines[1:]), "exactly one JSON line, from rank 0"
    return json.loads(lines[0][0])


def test_bench_control_flow_one_and_two_ranks(tmp_path):
    one = _run(tmp_path, 1, 29541)
    two = _run(tmp_path, 2, 29542)
    for line, n in ((one, 1), (two, 2)):
        assert line["n_gpus"] == n and line["steps"] == 1 and line["higher_is_better"] is True
        for key in ("metric", "value", "unit", "ms_per_step", "scaling", "dtype", "data", "config", "e2e", "roofline", "gpu_launches",
                    "roofline_kernels", "wall_ms_per_step", "counts_per_step", "stage_ms_per_step", "seed_hits_per_s", "gcells_per_s", "other_schedule"):
            assert key in line, key
        assert line["unit"] == "hits/s" and line["e2e"]["unit"] == "hits/s"           # the driver divides the two arms' e2e values
        assert line["value"] > 0 and line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0
        assert line["counts_per_step"]["segments_gathered"] == line["counts_per_step"]["hsps"]      # every rank's HSP table reached rank 0
        assert line["counts_per_step"]["alignments"] > 0 and line["counts_per_step"]["alignment_bytes_gathered"] > 0
        assert line["config"]["query_shards"] == n
        assert line["other_schedule"]["overlap"] is True and line["other_schedule"]["value"] > 0      # the default line is the one-after-the-other schedule
        assert line["config"]["chain"] == (n > 1)                                      # config3 on one rank, config4 (--chain) on several
    # the query is cut in two: every word of it is still scanned (windows across the seam aside), hits stay within 1 %
    h1, h2 = one["counts_per_step"]["raw_seed_hits"], two["counts_per_step"]["raw_seed_hits"]
    assert abs(h1 - h2) <= 0.01 * h1, (h1, h2)
