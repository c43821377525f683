"""CPU checks of bench.py's reference arm (`--impl reference`: the CPU restatement of the reference's path, timed on
the host cores): the JSON line carries the contract's keys, and under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e")


def _lines(cmd, env=None):
    e = dict(os.environ)
    # the reference arm must never map the product library: with SHPL_LIB pointing nowhere, importing sparse_pooling_b200
    # (whose __init__ loads libshpl.so) would raise
    e["SHPL_LIB"] = "/nonexistent/libshpl.so"
    e.update(env or {})
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, check=True, env=e).stdout
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    (d,) = _lines([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1"])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["higher_is_better"] is True and d["value"] > 0
    assert d["unit"] == "frames/s" and d["dtype"] == "f32" and "workload" in d["config"] and "model" not in d["config"]
    cb, e = d["cpu_baseline"], d["e2e"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_uses_every_host_core_whatever_the_launcher_exports():
    """torch.distributed.run exports OMP_NUM_THREADS=1 (round 1's N >= 2 baselines ran on one core): the arm sets the
    thread count itself, reports it, and marks the line rejected if the oracle still runs on fewer threads."""
    avail = len(os.sched_getaffinity(0))
    (d,) = _lines([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"], env={"OMP_NUM_THREADS": "1"})
    assert d["cpu_baseline"]["cores"] == avail and d["cpu_baseline"]["cores_available"] == avail and "rejected" not in d


def test_both_arms_print_the_same_config_object():
    """`config` says what the workload is and nothing about how one arm ran it: bench.config_dict is the only producer."""
    sys.path.insert(0, ROOT)
    import bench
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": config_dict(name, cfg,') == 3          # reference arm, avod GPU arm, pairs GPU arm
    for name, cfg in bench.configs().items():
        c = bench.config_dict(name, cfg, 1)
        assert c["baseline_config"] == name and "workload" in c and "model" not in c and "l2" in c
    (d,) = _lines([sys.executable, "bench.py", "--impl", "reference", "--config", "1", "--steps", "1", "--warmup", "0"])
    assert d["config"] == bench.config_dict("1", bench.configs()["1"], 1)


def test_reference_arm_under_torchrun_only_rank_zero_prints():
    lines = _lines([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                    "--master-port", "29571", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"])
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2
    assert lines[0]["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))        # not the 1 thread torchrun exports
