"""`--gpus=<n>` of the C front end: the query cut into n intervals, one process (and device) per interval, outputs gathered
in interval order.  The contract is the reference's own unit of independent work -- a query subrange `q.fa[a..b]` -- so the
output must equal, byte for byte, n reference runs on those subranges printed one after the other.  Checked here through
the oracle build of the front end (the device number means nothing to it); the multi-device run needs a multi-GPU box."""
import os
import subprocess

import pytest

from conftest import ORACLE_CLI, PRODUCT_CLI, REF_CLI, run_cli


def _query_length(path):
    return sum(len(l.strip()) for l in open(path) if not l.startswith(">"))


def _reference_by_intervals(t, q, n, opts):
    L, out = _query_length(q), ""
    for k in range(n):
        lo, hi = k * L // n, (k + 1) * L // n
        out += run_cli(REF_CLI, [t, f"{q}[{lo + 1}..{hi}]"] + opts)[0]
    return out


@pytest.mark.parametrize("n,opts", [(2, ["--format=general-"]), (3, []), (3, ["--format=maf", "--chain"]), (4, ["--nogapped", "--format=segments"]),
                                    (2, ["--strand=minus", "--format=axt", "--recoverseeds"])])
def test_gpus_option_equals_reference_on_the_same_intervals(synth, n, opts):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref not built")
    t, q = synth(300000)
    assert run_cli(ORACLE_CLI, [t, q, f"--gpus={n}"] + opts)[0] == _reference_by_intervals(t, q, n, opts)      # headers and command-line echoes included


def test_gpus_option_refusals(synth, tmp_path):
    t, q = synth(300000)
    two = tmp_path / "two.fa"
    two.write_text(">a\nACGTACGTACGTACGTACGTAAAACCCCGGGGTTTT\n>b\nACGTTTGACCAGTTGACAGTTTGACCA\n")
    for args, what in (([t, q + "[1..1000]", "--gpus=2"], "can't carry actions"), ([t, str(two), "--gpus=2"], "one sequence"),
                       ([t, "--self", "--gpus=2"], "needs a query file"), ([t, q, "--gpus=0"], "device count")):
        p = subprocess.run([ORACLE_CLI] + args, capture_output=True, text=True)
        assert p.returncode != 0 and what in p.stderr, (args, p.stderr[-300:])
    out = tmp_path / "o.txt"                                     # --output= belongs to the parent
    run_cli(ORACLE_CLI, [t, q, "--gpus=2", "--format=general-", f"--output={out}"])
    assert out.read_text() == run_cli(ORACLE_CLI, [t, q, "--gpus=2", "--format=general-"])[0]


@pytest.mark.gpu
def test_gpus_option_on_two_devices(synth):
    """needs two GPUs: skipped on the one-GPU boxes the suite normally runs on"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU here")
    t, q = synth(1000000)
    ref = REF_CLI if os.path.exists(REF_CLI) else ORACLE_CLI
    L, want = _query_length(q), ""
    for k in range(2):
        want += run_cli(ref, [t, f"{q}[{k * L // 2 + 1}..{(k + 1) * L // 2}]", "--format=general-"])[0]
    assert run_cli(PRODUCT_CLI, [t, q, "--gpus=2", "--format=general-"])[0] == want
