"""The host front end (product code: lastz_b200/csrc/host/*.c) built with AddressSanitizer + UndefinedBehaviorSanitizer
against the oracle library and driven through command lines that touch every reader and writer: any report fails."""
import glob
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT

HOST = os.path.join(ROOT, "lastz_b200", "csrc", "host")
CASES = [
    ["pseudocat.fa", "pseudopig.fa", "--format=maf", "--chain"],
    ["pseudocat.fa", "pseudopig.fa[multi]", "--format=softsam+eqx"],
    ["aglobin.2bit[multi]", "shorties.fa[multi]", "--format=general:name1,name2,cigarx,diff,text1,shingle", "K=2500", "--chain"],
    ["aglobin.2bit/human", "shorties.fq", "--yasra85", "--format=sam"],
    ["aglobin.2bit/human", "aglobin.2bit/cow", "K=top20%", "--format=axt"],
    ["aglobin.2bit/human", "--self", "--format=lav"],
    ["edge_target.fa", "edge_queries.fa[multi]", "--format=blastn"],
    ["names.fa[multi]", "names.fa", "--format=rdotplot+score", "K=2000"],
    ["aglobin.2bit/human", "aglobin.2bit/cow", "--identity=70", "--coverage=0.5", "--format=cigar"],
    ["pseudocat.fa", "pseudopig2.nib", "--format=gfa", "--nogapped"],
    ["pseudocat.fa", "pseudopig.fa", "--mismatch=2,25", "--chain", "--format=mapping"],
    ["aglobin.2bit/human", "shorties.fa", "--exact=20", "--anyornone", "--format=paf"],
    ["aglobin.2bit/human", "aglobin.2bit/cow", "--recoverseeds", "--format=maf-"],
    ["aglobin.2bit/human", "aglobin.2bit/cow", "--twins=0..40", "--seedqueue=500", "--format=general"],
    ["aglobin.2bit/human", "aglobin.2bit/cow", "--gpus=3", "--format=maf"],
    ["aglobin.2bit/human", "aglobin.2bit/cow", "--format=axt", "--segments=base_test.anchors.anchors"],
]


@pytest.fixture(scope="module")
def sanitized(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("asan") / "lastz_asan")
    src = sorted(glob.glob(os.path.join(HOST, "*.c"))) + [os.path.join(ROOT, "oracle", "lzb_oracle.c")]
    p = subprocess.run(["gcc", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-w", "-o", exe] + src +
                       ["-I" + os.path.join(ROOT, "include"), "-lm"], capture_output=True, text=True)
    if p.returncode != 0:
        pytest.skip("sanitizer build not available here: " + p.stderr[-300:])
    return exe


@pytest.mark.parametrize("args", CASES)
def test_front_end_is_clean_under_sanitizers(sanitized, args):
    argv = [os.path.join(GOLDEN, a) if not a.startswith("-") and "=" not in a.split("[")[0] else a for a in args]
    argv = [a.replace("--segments=", "--segments=" + GOLDEN + "/") for a in argv]
    p = subprocess.run([sanitized] + argv, capture_output=True, text=True, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    assert "AddressSanitizer" not in p.stderr and "runtime error" not in p.stderr, p.stderr[-2000:]
    assert p.returncode == 0, p.stderr[-500:]
