"""CPU suite: the parts of bench.py's contract that run without a GPU -- the reference arm's JSON line and the
algorithmic-bytes formula of SURVEY.md 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    sys.path.insert(0, ROOT)
    import bench
    return bench


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` times the reference's own CPU path (oracle/_ref, or the port when the reference
    was not built) and prints one JSON line with the arm's keys."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--frames", "2", "--classes", "2"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == _bench().METRIC and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["frames_per_gpu"] == 2 and d["config"]["classes"] == 2
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_stage_bytes_sum_to_the_survey_formula():
    """bench.stage_bytes splits SURVEY 8(d)'s B_frame = 12P + 12KP + 24(d+1)P + 8(d+1)M + (d+1)(8M + 16KM) + 8KM."""
    bench = _bench()
    for P, d, K, M in ((224 * 224, 5, 2, 56141.0), (224 * 224, 5, 10, 3070.0), (448 * 448, 3, 2, 735.6), (1, 1, 1, 2.0)):
        want = 12 * P + 12 * K * P + 24 * (d + 1) * P + 8 * (d + 1) * M + (d + 1) * (8 * M + 16 * K * M) + 8 * K * M
        assert abs(bench.frame_bytes(P, d, K, M) - want) < 1e-6 * want
    # the worked example of the survey: 224^2, d = 5, K = 2, noise (M = 56 141) -> 26.1 MB per frame
    assert abs(bench.frame_bytes(224 * 224, 5, 2, 56141.0) / 1e6 - 26.1) < 0.1
