"""CPU: the bench.py output contract.  The reference arm (`--impl reference`, the reference's own functions on the host cores) prints
exactly one JSON line with the agreed keys; the committed N=1 line of our arm (profiles/r1_bench_n1.json, produced on a
B200) carries roofline / cpu_baseline / e2e / clocks / gpu_launches in the agreed shape."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["value"] > 0 and d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("synthetic Jaccard top-K")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "retrieval_data_annotation.bytecode")):
        assert cb["kind"] == "reference"          # the unmodified reference functions are what is timed
    assert cb["per_core"]["value"] > 0 and d["extrapolated"] is True and d["sample_pairs_per_step"] > 0
    import bench
    ns = bench.parse_args.__globals__["argparse"].Namespace(pool=bench.POOL_N, queries=bench.QUERY_N, mean_set=1.0 / 0.45)
    assert d["config"] == bench.workload_config(ns, 1)          # both arms print the same config dict
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_committed_bench_line_has_the_contract_shape():
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_n1.json")))
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks", "verified"} <= set(d)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["data"] == "synthetic"
    assert d["verified"] is True and all(v is True for k, v in d["verified_what"].items() if k != "how")
    rf = d["roofline"]
    assert rf["bound"] in ("hbm", "tensor") and rf["unit"] in ("GB/s", "TFLOP/s")
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and rf["frac"] > 0
    # the numerator is SURVEY 8(d)'s compulsory bytes of the bitset representation; the postings kernel's own bytes and
    # the measured DRAM traffic are reported next to it (they are far smaller: the bitsets are never read)
    own = rf["own_representation"]
    assert own["bytes"] > 0 and own["frac_hbm"] < rf["frac"] and own["postings_visited"] > 0
    assert rf["traffic"] is None or 0 < rf["traffic"] < rf["algorithmic_bytes"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"] and e["pipelined"]["value"] < d["value"]          # copies included: never faster
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"]
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
    assert "model" not in d["config"] and d["config"]["workload"] and "L2 flushed" in d["config"]["l2"]
    bp = d["bitset_path"]["roofline"]
    assert bp["bound"] == "hbm" and 0 < bp["frac"] <= 1.0
    dn = d["dense"]
    assert dn["roofline"]["bound"] == "tensor" and dn["e2e"]["value"] > 0 and dn["gpu_launches"] > 0 and dn["verified"] is True
    assert "BF16X3" in dn["config"]["precision"] and 0 < dn["roofline"]["frac"] <= 1.0
    assert dn["variants"]["bf16_single_pass"]["value"] > dn["value"]                 # narrower arithmetic is faster, and labelled


def test_committed_scaling_lines_are_weak_and_verified():
    """Two measurement series are kept: the final one (head / big kernels: N = 1, 2 and the Jaccard line at N = 8) and the
    earlier register-kernel series of this round (N = 1, 2, 4, 8 with the fixed-workload curves).  Each series is
    compared with its own N = 1 line."""
    prof = os.path.join(ROOT, "profiles")
    base = json.load(open(os.path.join(prof, "r2_bench_n1.json")))
    for n, name in ((2, "r2_bench_n2.json"), (8, "r2_bench_n8_jaccard.json")):
        d = json.load(open(os.path.join(prof, name)))
        assert d["n_gpus"] == n and d["scaling"] == "weak" and d["verified"] is True
        assert "query-sharded" in d["config"]["parallelism"]
        assert d["value"] / (n * base["value"]) > 0.9                                # weak-scaling efficiency
        assert d["e2e"]["value"] > base["e2e"]["value"]
    st = json.load(open(os.path.join(prof, "r2_bench_n2.json")))["strong"]
    assert st["query_sharded"]["value"] > 0 and st["pool_sharded"]["value"] > 0
    old = json.load(open(os.path.join(prof, "r2_bench_regkernel_n1.json")))
    assert old["value"] < base["value"]                                              # the head kernel is the faster one
    for n in (2, 4, 8):
        d = json.load(open(os.path.join(prof, f"r2_bench_regkernel_n{n}.json")))
        assert d["n_gpus"] == n and d["scaling"] == "weak" and d["verified"] is True
        assert "query-sharded" in d["config"]["parallelism"]
        assert d["value"] / (n * old["value"]) > 0.9
        st = d["strong"]
        assert st["query_sharded"]["value"] > 0 and st["pool_sharded"]["value"] > 0
