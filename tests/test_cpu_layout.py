"""Repository-level rules that can be checked without a GPU: the product path never touches the
oracle (test infrastructure) or the reference checkout, and the reference arm of bench.py answers."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT = ["ti_sph_b200", "core", "utils", "main_3d.py", "main.py", "demo.py"]


def product_files():
    for entry in PRODUCT:
        path = os.path.join(ROOT, entry)
        if os.path.isfile(path):
            yield path
            continue
        for base, _, names in os.walk(path):
            for n in names:
                if n.endswith((".py", ".cu", ".cuh", ".h")):
                    yield os.path.join(base, n)


def test_product_path_never_imports_the_oracle_or_reads_the_reference():
    bad = []
    for f in product_files():
        text = open(f, encoding="utf-8").read()
        if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "liboracle" in text:
            bad.append((f, "oracle"))
        if "/root/reference" in text:
            bad.append((f, "/root/reference"))
    assert not bad, bad


def test_gpu_side_entry_points_do_not_read_the_reference_checkout():
    for f in ["bench.py", "__graft_entry__.py"] + [os.path.join("tests", n) for n in os.listdir(os.path.join(ROOT, "tests"))
                                                     if n.startswith("test_gpu")]:
        assert "/root/reference" not in open(os.path.join(ROOT, f), encoding="utf-8").read(), f


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "3", "--ref-particles", "20000", "--workload", "C3"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "particle_updates_per_sec"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0


def test_every_kernel_source_is_in_the_build_hash():
    from ti_sph_b200 import build
    listed = set(build.SOURCES + build.HEADERS)
    present = {n for n in os.listdir(build.CSRC) if n.endswith((".cu", ".cuh"))}
    assert present == listed, present ^ listed
