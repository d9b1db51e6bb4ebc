"""bench.py's reference arm runs without a GPU and prints the contract's JSON line: the unmodified reference model from
oracle/_ref when oracle/build_ref.py has built it (kind "reference"), else the oracle port (kind "port")."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"] + extra,
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=e)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_reference_arm_prints_the_contract_line():
    line = _run(["--batch", "64"], env={"MFM_NO_REF": "1"})
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == dict(value=line["value"], unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert line["gpu_launches"] == 0 and "workload" in line["config"] and line["scaling"] == "weak"


def test_reference_arm_uses_the_unmodified_reference_when_built():
    sys.path.insert(0, ROOT)
    from oracle import build_ref, ref_runner
    if not os.path.exists(build_ref.REF_SRC) and not ref_runner.available():
        pytest.skip("no reference tree and no prebuilt oracle/_ref here")
    build_ref.build(verbose=False)
    line = _run(["--batch", "32", "--config", "iemocap"])
    assert line["cpu_baseline"]["kind"] == "reference" and line["value"] > 0
    assert line["config"]["name"] == "iemocap" and line["config"]["seq_len"] == 20


def test_workloads_and_strong_scaling_arguments():
    line = _run(["--config", "mosei", "--batch", "16"], env={"MFM_NO_REF": "1"})
    assert line["scaling"] == "strong" and line["config"]["global_batch"] == 16 and line["config"]["seq_len"] == 50
    line = _run(["--config", "pom", "--batch", "8"], env={"MFM_NO_REF": "1"})
    assert line["scaling"] == "weak" and line["config"]["seq_len"] == 100


def test_roofline_traffic_lookup_is_deterministic():
    """bench.py names the GEMM shape whose ncu DRAM traffic it reports from the workload's DIMENSIONS; the committed ncu
    extract must hold that key for the headline workload (round 1 printed traffic = null on the driver's box)."""
    sys.path.insert(0, ROOT)
    import bench
    from factorized_b200.configs import WORKLOADS, best_acc_configs
    from factorized_b200.engine import Dims
    wl = WORKLOADS["mosi"]
    dm = Dims(best_acc_configs(), wl["T"], wl["batch"])
    key, nbytes = bench.top_gemm_shape(dm)
    assert key == "nt 40960x400x128" and nbytes == 4.0 * (40960 * 128 + 400 * 128 + 40960 * 400)
    nj = json.load(open(os.path.join(ROOT, "profiles", "r2_gemm_tcp_ncu.json")))
    assert key in nj and nj[key]["dram_bytes"] > 0
