"""bench.py on a CPU-only box: the byte model matches BASELINE.md 3, and the reference arm (the threaded C
restatement of the reference's CPU algorithm, the one place outside tests/ where the oracle may run) prints one
JSON line with every key the measurement contract names."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench        # noqa: E402


def test_vcycle_byte_model_matches_baseline_md():
    # Tri r=4: 702 B, Tri r=7: 670 B, Tet r=4: 619 B, Tet r=5: 605 B per finest stored DOF
    assert round(bench.vcycle_bytes_per_dof(2, 5)) == 702
    assert round(bench.vcycle_bytes_per_dof(2, 8)) == 670
    assert round(bench.vcycle_bytes_per_dof(3, 5)) == 619
    assert round(bench.vcycle_bytes_per_dof(3, 6)) == 605


def test_workloads_are_the_baseline_configs():
    assert bench.WORKLOADS["C1"]["dim"] == 2 and bench.WORKLOADS["C1"]["levels"] == 5 and bench.WORKLOADS["C1"]["c"] == 48
    assert bench.WORKLOADS["C2"]["dim"] == 2 and bench.WORKLOADS["C2"]["levels"] == 8
    assert bench.WORKLOADS["C3"]["dim"] == 3 and bench.WORKLOADS["C3"]["levels"] == 5 and bench.WORKLOADS["C3"]["c"] == 20
    assert bench.WORKLOADS["C4"]["dim"] == 3 and bench.WORKLOADS["C4"]["levels"] == 6 and bench.WORKLOADS["C4"]["c"] == 32
    assert bench.WORKLOADS["C2"]["c"] == 256
    # C4's coarse problem fits the 32-bit potrf + potri path (fewer than 46 340 interior base nodes); C2 at c = 256
    # (65 025) takes the 64-bit one
    assert (bench.WORKLOADS["C4"]["c"] - 1) ** 3 < 46340 <= (bench.WORKLOADS["C2"]["c"] - 1) ** 2


def test_one_workload_at_every_gpu_count():
    """The scaling curve is one problem: the primary workload does not depend on WORLD_SIZE."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'name = "C4" if args.workload == "auto" else args.workload' in src
    assert "parity_vs_single_gpu" in src


@pytest.mark.slow
def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C3", "--cells", "3",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "GDOF/s" and line["dtype"] == "f64" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
