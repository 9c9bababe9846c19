"""bench.py contract pieces that can be checked without a GPU: the result is the ONLY thing on stdout (libraries such as NCCL print
banners on file descriptor 1 under torchrun), and the reference arm answers for models it has no CPU port for."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_result_line_is_alone_on_stdout():
    code = ("import bench, os; bench.reserve_stdout(); os.write(1, b'NCCL version x.y\\n'); print('python noise'); "
            "bench.emit({'metric': 'm', 'value': 1.0})")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=120)
    assert out.returncode == 0
    assert out.stdout.count("\n") == 1 and json.loads(out.stdout) == {"metric": "m", "value": 1.0}
    assert "NCCL version" in out.stderr and "python noise" in out.stderr


def test_reference_arm_declines_models_without_a_cpu_port():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--model", "qwen2.5-vl-7b"], capture_output=True, text=True,
                         cwd=ROOT, timeout=300)
    assert out.returncode == 0
    line = json.loads(out.stdout)
    assert line["impl"] == "reference" and "unavailable" in line


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2"], capture_output=True, text=True, cwd=ROOT,
                         timeout=300, env=env)
    assert out.returncode == 0 and out.stdout == ""
