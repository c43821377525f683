"""CPU checks of bench.py's reference arm (`--impl reference`: the CPU restatement of the reference's path, timed on
the host cores): the JSON line carries the contract's keys, and under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e")


def _lines(cmd):
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, check=True).stdout
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


def test_reference_arm_under_torchrun_only_rank_zero_prints():
    lines = _lines([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                    "--master-port", "29571", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"])
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2
