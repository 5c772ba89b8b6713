"""bench.py contract pieces that can be checked without a GPU: the reference arm (`--impl reference`, the CPU path of the
reference through the oracle port) prints ONE JSON line with the keys the driver reads, on the GPU arm's config object."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_json_line():
    # configs[1] (N = 100k) keeps this CPU test short; the default workload is the north-star size, checked below
    d = _run("--impl", "reference", "--workload", "c2", "--steps", "1", "--warmup", "0")
    assert d["impl"] == "reference" and d["metric"].startswith("queries/sec") and d["unit"] == "queries/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    # oracle/_ref (the verbatim copy of the reference file, oracle/build_ref.py) is there wherever build() has run
    have_ref = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "eval", "matrix_approx_zeshel.py"))
    assert cb["kind"] == ("reference" if have_ref else "port")
    assert cb["cores"] == os.cpu_count() and cb["value"] == d["value"] and "sample" in cb
    cfg = d["config"]
    assert cfg["n_items"] == 100000 and cfg["k_i"] == 500 and cfg["batch"] == 4096 and cfg["top_k"] == 100
    assert "workload" in cfg and cfg["workload"].startswith("c2")


def test_default_workload_is_the_north_star_size_and_items_are_sharded():
    sys.path.insert(0, ROOT)
    import argparse
    import importlib
    fd = os.dup(1)                                   # importing bench redirects fd 1 to stderr (one-JSON-line rule): undo it
    try:
        bench = importlib.import_module("bench")
    finally:
        os.dup2(fd, 1)
        os.close(fd)
    assert bench.WORKLOADS["n1m"] == (1_000_000, 500, 2000, 4096, 100)
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'ap.add_argument("--workload", default="n1m"' in src and 'shard = args.shard or "items"' in src
    ns = argparse.Namespace(workload="n1m", precision="f32r", exchange="p2p")
    one, eight = bench.workload_config(ns, 1, "items"), bench.workload_config(ns, 8, "items")
    assert one["n_items"] == 1_000_000 and one["workload"].startswith("n1m") and "whole index" in one["parallelism"]
    assert "items sharded over 8 ranks" in eight["parallelism"] and "NVLink" in eight["parallelism"]


def test_reference_arm_other_ranks_print_nothing():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""
