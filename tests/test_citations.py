"""Every `file:line` citation of the reference in the C ABI header, the product sources and the oracle points
at a file that exists in /root/reference and at lines inside it.  Runs only where the reference tree is mounted
(this container); on the GPU box it is skipped."""
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
CITE = re.compile(r"(?<![\w/.-])((?:[\w.-]+/)*[\w.-]+\.(?:cpp|h|hpp|txt|md)|bin/run-suite):(\d+)(?:-(\d+))?")
# citations of this repository's own files (not the reference's)
OWN_PREFIXES = ("sb_", "tests/", "profiles/", "oracle/", "sycl-bench_b200/", "DESIGN", "INTEGRATION", "BASELINE", "SURVEY", "README")


def _sources():
    pats = ["include/*.h", "sycl-bench_b200/csrc/*.cu*", "sycl-bench_b200/harness/*.h", "sycl-bench_b200/pattern/*", "sycl-bench_b200/sycl2020/*",
            "sycl-bench_b200/single-kernel/*", "sycl-bench_b200/micro/*", "sycl-bench_b200/run-suite", "sycl-bench_b200/*.py", "oracle/*.c", "oracle/*.py",
            "oracle/shim/sycl/sycl.hpp"]
    out = []
    for p in pats:
        out += [f for f in glob.glob(os.path.join(ROOT, p)) if os.path.isfile(f)]
    return sorted(out)


def _resolve(name):
    for cand in (name, os.path.join("include", name), os.path.join("pattern", name), os.path.join("polybench/common", name)):
        path = os.path.join(REF, cand)
        if os.path.isfile(path):
            return path
    hits = glob.glob(os.path.join(REF, "**", os.path.basename(name)), recursive=True)
    return hits[0] if len(hits) == 1 else None


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_reference_citations_resolve():
    checked, bad = 0, []
    lengths = {}
    for src in _sources():
        text = open(src, errors="replace").read()
        for m in CITE.finditer(text):
            name, lo, hi = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            if name.startswith(OWN_PREFIXES) or os.path.basename(name).startswith("sb_"):
                continue
            path = _resolve(name)
            if path is None:
                bad.append((os.path.relpath(src, ROOT), m.group(0), "no such reference file"))
                continue
            if path not in lengths:
                lengths[path] = sum(1 for _ in open(path, errors="replace"))
            checked += 1
            if not (1 <= lo <= hi <= lengths[path]):
                bad.append((os.path.relpath(src, ROOT), m.group(0), f"file has {lengths[path]} lines"))
    assert checked > 80, checked
    assert not bad, bad[:20]
