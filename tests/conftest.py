import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_CLI = os.path.join(ROOT, "oracle", "lastz_oracle")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "lastz")
PRODUCT_CLI = os.path.join(ROOT, "lastz_b200", "csrc", "lastz_b200")
GEN_SYNTH = os.path.join(ROOT, "tools", "gen_synth")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    # the checker (oracle + synthetic generator) is plain gcc; build it on demand
    if not (os.path.exists(ORACLE_CLI) and os.path.exists(os.path.join(ROOT, "oracle", "liblzb_oracle.so"))):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liblzb_oracle.so", "lastz_oracle"], check=True)
    if not os.path.exists(GEN_SYNTH):
        subprocess.run(["gcc", "-O2", "-o", GEN_SYNTH, os.path.join(ROOT, "tools", "gen_synth.c")], check=True)
    if not os.path.exists(os.path.join(ROOT, "lastz_b200", "csrc", "liblzb_host.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "lastz_b200", "csrc"), "liblzb_host.so"], check=True)


def run_cli(cli, args, cwd=None):
    """stdout of one command-line run; stderr is returned too (the reference warns there)."""
    p = subprocess.run([cli] + list(args), cwd=cwd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"{cli} {' '.join(args)} failed ({p.returncode}): {p.stderr[-2000:]}")
    return p.stdout, p.stderr


def lav_body(text):
    """LAV with the d-stanza command line dropped (tools/lav_compare.py:41-78 ignores it too)."""
    out, in_d, ix = [], False, 0
    for line in text.splitlines():
        s = line.rstrip()
        if s == "d {":
            in_d, ix = True, 0
            out.append(s)
            continue
        if in_d:
            ix += 1
            if s == "}":
                in_d = False
            elif ix == 1:
                continue
        out.append(s)
    return out


@pytest.fixture(scope="session")
def synth(tmp_path_factory):
    """synthetic pairs by size (SURVEY.md 8d generator), cached per session"""
    d = tmp_path_factory.mktemp("synth")
    made = {}

    def get(n, seed=20260925):
        key = (n, seed)
        if key not in made:
            t, q = str(d / f"t{n}_{seed}.fa"), str(d / f"q{n}_{seed}.fa")
            subprocess.run([GEN_SYNTH, str(n), str(seed), t, q], check=True)
            made[key] = (t, q)
        return made[key]
    return get
