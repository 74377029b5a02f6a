"""The one-lease GPU measurement plan (tools/gpu_first_call.sh) must stay runnable: valid bash, every test file and every
bench.py flag it names exists."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_first_call_script_is_consistent():
    path = os.path.join(ROOT, "tools", "gpu_first_call.sh")
    assert subprocess.run(["bash", "-n", path]).returncode == 0
    text = open(path).read()
    for f in re.findall(r"tests/[\w/]+\.py", text):
        assert os.path.exists(os.path.join(ROOT, f)), f
    bench = open(os.path.join(ROOT, "bench.py")).read()
    for flag in set(re.findall(r"(--[a-z][a-z0-9-]+)", " ".join(ln for ln in text.splitlines() if "bench.py" in ln))):
        if flag in ("--metrics", "--clock-control", "--csv", "--log-file", "--set", "--import-source"):
            continue   # ncu's own
        assert f'"{flag}"' in bench, flag
    gpu_files = sorted(f for f in os.listdir(os.path.join(ROOT, "tests")) if f.startswith("test_") and f.endswith(".py")
                       and re.search(r"^pytestmark = pytest\.mark\.gpu", open(os.path.join(ROOT, "tests", f)).read(), re.M))
    for f in gpu_files:
        assert f"tests/{f}" in text, f"{f} is a GPU test file the script does not run"
