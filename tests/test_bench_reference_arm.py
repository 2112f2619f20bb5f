"""`bench.py --impl reference` is the arm the driver times beside the CUDA arm: the oracle (the C port of the reference's CPU path;
the Rust crate cannot be built in this image) over the same frames, no GPU involved.  It must print ONE JSON line with the contract's
keys, `impl: "reference"`, the same `metric` / `config.workload` as the CUDA arm, and `e2e` equal to its own value; ranks other than 0
print nothing and exit 0."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", *args], capture_output=True, text=True, env=env,
                          timeout=600, check=True)


def test_reference_arm_prints_the_contract_line():
    out = _run({}, "--workload", "C3", "--batch", "4", "--steps", "1", "--warmup", "0").stdout.strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    assert line["impl"] == "reference" and line["metric"] == "frames_per_sec_1080p_batch256" and line["unit"] == "frames/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert line["config"]["workload"].startswith("C3: 1920x1080 RGB8 x 256 frames per GPU")
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 0 and line["markers_per_step"] >= 4 * 15 and line["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2", "--workload", "C3", "--batch", "4", "--steps", "1", "--warmup", "0")
    assert r.stdout.strip() == ""


def test_cuda_arm_refuses_to_run_without_a_gpu():
    """The product path has no CPU fallback: without a CUDA device the CUDA arm says so and fails (it never turns to the oracle)."""
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
