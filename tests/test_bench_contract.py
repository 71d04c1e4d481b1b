"""bench.py JSON contract of the reference arm (runs on CPU with a toy geometry: PCL_BENCH_TINY=1)."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _run(extra):
    env = dict(os.environ, PCL_BENCH_TINY="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1"] + extra, capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_json_line():
    d = _run([])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["value"] > 0 and "workload" in d["config"]


def test_reference_arm_bank_workload():
    d = _run(["--workload", "s2"])
    assert d["impl"] == "reference" and "memory bank" in d["config"]["workload"]


def test_graph_arm_protocol_never_blocks_the_parent(tmp_path):
    """bench.py's CUDA-graph arm runs in a child process per rank (READY / GO / one JSON line).  Whatever the child does —
    refuses (no GPU here), answers after unrelated output, dies, or stays silent — the parent gets a definite answer."""
    import subprocess
    import sys
    import types
    sys.path.insert(0, ROOT)
    import bench

    def arm_for(cmd):
        arm = bench.GraphArm.__new__(bench.GraphArm)
        arm.result, arm._buf = None, b""
        arm.proc = subprocess.Popen(cmd, stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
        return arm

    args = types.SimpleNamespace(steps=3, warmup=3, precision="bf16", workload="s1", scaling="weak")
    real = bench.GraphArm(args, 0, 0)                       # the real child: no CUDA device on this box -> a refusal, not READY
    assert real.wait_ready(180) is False and real.result["ok"] is False and real.result["why"]
    real.close()
    fake = tmp_path / "child.py"
    fake.write_text('import sys, json\nprint("library noise")\nprint("READY", flush=True)\n'
                    'assert sys.stdin.readline().strip() == "GO"\n'
                    'print(json.dumps({"ok": True, "ms_per_step": 0.1, "e2e_s_per_step": 0.007, "finite": True}), flush=True)\n')
    arm = arm_for([sys.executable, str(fake)])
    assert arm.wait_ready(60) is True
    assert arm.go(60) == {"ok": True, "ms_per_step": 0.1, "e2e_s_per_step": 0.007, "finite": True}
    arm.close()
    assert arm.proc.poll() == 0
    arm = arm_for([sys.executable, "-c", "import sys; sys.exit(3)"])
    assert arm.wait_ready(60) is False and arm.result["ok"] is False
    arm.close()
    arm = arm_for([sys.executable, "-c", "import time; time.sleep(120)"])
    assert arm.wait_ready(1.0) is False
    arm.close()
    assert arm.proc.poll() is not None                      # the silent child was terminated (exact pid)
