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


_FAKE_OK = ('import sys, json\nprint("READY", flush=True)\nassert sys.stdin.readline().strip() == "GO"\n'
            'print(json.dumps({"ok": True, "ms_per_step": %s, "e2e_s_per_step": 0.008, "host_enqueue_ms_per_step": 0.01, '
            '"e2e_steps": 50, "clocks": None, "finite": True}), flush=True)\n')


def _fake_arm_factory(bench, script_for_rank):
    import subprocess

    class FakeArm(bench.GraphArm):
        def __init__(self, args, rank, local):
            self.result, self._buf = None, b""
            self.proc = subprocess.Popen([sys.executable, script_for_rank(rank)], stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL)
    return FakeArm


def test_graph_arm_reduction_single_rank(tmp_path, monkeypatch):
    import types
    import torch
    sys.path.insert(0, ROOT)
    import bench
    ok, bad = tmp_path / "ok.py", tmp_path / "bad.py"
    ok.write_text(_FAKE_OK % "0.1")
    bad.write_text('import json\nprint(json.dumps({"ok": False, "why": "replay 0 differs from the eager step"}), flush=True)\n')
    args = types.SimpleNamespace(steps=200, warmup=10, precision="bf16", workload="s1", scaling="weak")
    cfg = {"B": 8}
    monkeypatch.setattr(bench, "GraphArm", _fake_arm_factory(bench, lambda r: str(ok)))
    g = bench.graph_arm_measure(args, cfg, 0, 1, torch.device("cpu"), lambda: None)
    assert g["ok"] and abs(g["ms_per_step"] - 0.1) < 1e-12 and abs(g["value"] - 8 / 0.1e-3) < 1e-6
    assert abs(g["e2e_value"] - 8 / 0.008) < 1e-6 and g["e2e_steps"] == 50
    monkeypatch.setattr(bench, "GraphArm", _fake_arm_factory(bench, lambda r: str(bad)))
    g = bench.graph_arm_measure(args, cfg, 0, 1, torch.device("cpu"), lambda: None)
    assert g == {"ok": False, "why": "replay 0 differs from the eager step"}


def _graph_arm_worker(rank, world, port, out_dir, scripts):
    import types
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bench.GraphArm = _fake_arm_factory(bench, lambda r: scripts[r])
    args = types.SimpleNamespace(steps=200, warmup=10, precision="bf16", workload="s1", scaling="weak")
    g = bench.graph_arm_measure(args, {"B": 8}, rank, world, torch.device("cpu"), dist.barrier)
    with open(os.path.join(out_dir, f"g{rank}.json"), "w") as f:
        json.dump(g, f)
    dist.barrier()
    dist.destroy_process_group()


def test_graph_arm_reduction_two_ranks(tmp_path):
    """Max over ranks when both children succeed; when one rank's child refuses, BOTH ranks fall back (and nobody hangs)."""
    import socket
    import torch.multiprocessing as mp
    fast, slow, bad = tmp_path / "fast.py", tmp_path / "slow.py", tmp_path / "bad.py"
    fast.write_text(_FAKE_OK % "0.10")
    slow.write_text(_FAKE_OK % "0.13")
    bad.write_text('import json\nprint(json.dumps({"ok": False, "why": "CUDA error"}), flush=True)\n')
    for scripts, expect_ok in (([str(fast), str(slow)], True), ([str(fast), str(bad)], False)):
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        mp.spawn(_graph_arm_worker, args=(2, port, str(tmp_path), scripts), nprocs=2, join=True)
        g0, g1 = (json.load(open(tmp_path / f"g{r}.json")) for r in (0, 1))
        assert g0["ok"] == g1["ok"] == expect_ok
        if expect_ok:
            assert abs(g0["ms_per_step"] - 0.13) < 1e-12 and g0["value"] == g1["value"]
            assert abs(g0["value"] - 2 * 8 / 0.13e-3) < 1e-6            # max over ranks sets the whole-job throughput


def _run_engine_on_emulator(monkeypatch, fake_graph_ms=None):
    """bench.run_engine end to end on the CPU: the engine's kernels on the host-fiber emulator (tests/emu), CUDA events /
    pinned memory stubbed, toy geometry.  Exercises everything between argument parsing and the JSON line."""
    import types
    import torch
    sys.path.insert(0, ROOT)
    import bench
    import emu_harness
    emu_harness.use_emulation(monkeypatch)

    class Ev:
        def __init__(self, **k):
            pass

        def record(self):
            pass

        def elapsed_time(self, other):
            return 2.0

    monkeypatch.setattr(torch.cuda, "Event", Ev)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    if fake_graph_ms is not None:
        monkeypatch.setattr(bench, "graph_arm_measure", lambda *a, **k: {
            "ok": True, "ms_per_step": fake_graph_ms, "value": 2 / (fake_graph_ms / 1e3), "e2e_value": 123.0, "e2e_steps": 50,
            "host_enqueue_ms_per_step": 0.01, "clocks": {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 4},
            "check": "stub"})
    orig_stage = bench.stage_timings
    monkeypatch.setattr(bench, "stage_timings", lambda *a, **k: orig_stage(*a, iters=2))        # the default 20 only averages
    args = types.SimpleNamespace(steps=4, warmup=3, precision="fp32", workload="s1", scaling="weak", graph=False,
                                 no_graph_arm=False, no_cpu_baseline=False, gpus=1)
    cfg = dict(bench.S1)
    cfg.update(B=2, D=32, h=16, w=16, K=5, stride=2, block=8, max_samples=32, max_views=4)
    res = bench.run_engine(args, cfg, False, 0, 1, torch.device("cpu"))
    return json.loads(json.dumps(res))                     # must be JSON-serialisable as is


REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline")


def test_engine_arm_line_on_the_emulator_eager_fallback(monkeypatch):
    res = _run_engine_on_emulator(monkeypatch)                      # the real graph child refuses (no GPU here) -> eager line
    for k in REQUIRED:
        assert k in res, k
    assert res["impl"] == "engine" and res["cuda_graph"] is False and res["graph_arm"]["ok"] is False
    assert res["value"] == res["eager"]["value"] and res["config"]["step"].startswith("eager")
    assert res["ms_per_step"] == 0.5 and res["value"] == 2 * 4 / 2e-3          # 4 steps in the stubbed 2 ms
    assert res["roofline"]["bound"] == "hbm" and res["cpu_baseline"]["kind"] == "port"
    assert res["train_iter"] and "error" not in res["train_iter"] and res["train_iter"]["finite_loss"] is True
    assert set(res["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}


def test_engine_arm_line_on_the_emulator_graph_headline(monkeypatch):
    monkeypatch.setenv("PCL_BENCH_NO_TRAIN_ITER", "1")              # covered by the test above
    res = _run_engine_on_emulator(monkeypatch, fake_graph_ms=0.2)   # a verified, faster graph arm becomes the headline
    assert res["cuda_graph"] is True and res["ms_per_step"] == 0.2 and res["value"] == 2 / 0.2e-3
    assert res["eager"]["ms_per_step"] == 0.5 and res["e2e"]["value"] == 123.0 and res["e2e"]["steps"] == 50
    assert res["config"]["step"].startswith("one CUDA-graph replay") and res["clocks"]["samples"] == 4
