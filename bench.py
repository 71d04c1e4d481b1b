#!/usr/bin/env python
"""bench.py — contrast-loss fwd+bwd throughput on synthetic Cityscapes-shaped batches (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference] [--workload s1|s2]

Workload s1 = BASELINE configs[1]: HRNet-W48 pixel-contrast (no bank), 1024x512 input -> 256x128 embedding,
19 classes, batch 8 PER GPU (weak scaling; the no-bank loss has no cross-rank exchange).
Workload s2 = configs[2] shape per rank (bank 19x(5000+5000)x256, one image per rank, bank merge allgather).
A "step" = one pass of the hot path over one batch: label/argmax/key scan, anchor selection + gather, InfoNCE
forward, InfoNCE backward and the dense embedding-gradient write (+ bank enqueue for s2).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

S1 = dict(B=8, D=256, h=128, w=256, K=19, stride=4, block=32, T=0.1, bT=0.07, max_samples=1024, max_views=100)
S2 = dict(B=1, D=256, h=128, w=256, K=19, stride=4, block=32, T=0.07, bT=0.07, max_samples=1024, max_views=100,
          M=5000, F=10, net_stride=4)
# BASELINE configs[3], opt-in: ResNet-101 DeepLab-V3 + region memory on COCO-Stuff 520x520, 171 classes,
# 2 images per rank; the dilated-8 backbone yields a 66x66 embedding (nearest label index floor(dst*520/66), not ::8) and
# the bank's labels[:, ::8, ::8] grid (65x65) is flat-index-misaligned with the 4356 feature columns (Q6)
S3 = dict(B=2, D=256, h=66, w=66, K=171, stride=8, block=40, himg=520, wimg=520, T=0.07, bT=0.07, max_samples=1024,
          max_views=100, M=5000, F=10, net_stride=8)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
    (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from a .ncu-rep); None if not captured."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get(kernel, {}).get("dram_bytes")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        time.sleep(0.12)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        self.lines = [l for l in out.splitlines() if l.strip()]

    def summary(self):
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = max(smax, float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": smax or None, "reasons": sorted(reasons), "samples": len(sm)}


def pin_to_gpu_local_cpus(dev_index: int):
    """Multi-rank runs: bind this process to the CPUs NVML reports as local to its GPU (NUMA-aware launch path; the
    step is host-launch-bound, and ranks scheduled on the far socket launched ~45 % slower at N=8).  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(dev_index).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def make_inputs(cfg, seed, device=None, bank=False):
    from contrastiveseg_b200.synth import make_bank, make_contrast_batch
    d = make_contrast_batch(B=cfg["B"], D=cfg["D"], h=cfg["h"], w=cfg["w"], num_classes=cfg["K"],
                            img_stride=cfg["stride"], block=cfg["block"], seed=seed, himg=cfg.get("himg"),
                            wimg=cfg.get("wimg"))
    out = dict(embed=d["embed"], seg=d["seg"], target=d["target"])
    if bank:
        out.update(make_bank(cfg["K"], cfg["M"], cfg["D"], seed + 1000))
    if device is not None:
        out = {k: v.to(device) for k, v in out.items()}
    return out


def engine_configer(cfg, bank=False, precision="bf16"):
    import contrastiveseg_b200 as cs
    d = {"data": {"num_classes": cfg["K"]}, "network": {"stride": cfg.get("net_stride", 4)},
         "loss": {"params": {"ce_ignore_index": -1, "ce_reduction": "elementwise_mean"}},
         "contrast": {"temperature": cfg["T"], "base_temperature": cfg["bT"], "max_samples": cfg["max_samples"],
                      "max_views": cfg["max_views"], "loss_weight": 0.1, "use_rmi": False, "rng": "device",
                      "precision": precision}}
    if bank:
        # bank sweeps run on the tcgen05 path (bf16 operands through the bank shadow)
        d["contrast"].update(with_memory=True, memory_size=cfg["M"], pixel_update_freq=cfg["F"])
    return cs.Configer(d)


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores (the reference is Python and cannot
# travel to the GPU box; oracle/ref_port.py restates it op for op and is pinned to it by golden vectors)
# ---------------------------------------------------------------------------------------------------
def cpu_port_step(inp, cfg, bank=False):
    from oracle import ref_port as P
    embed = inp["embed"].clone().requires_grad_(True)
    predict = inp["seg"].argmax(1)
    queue = torch.cat((inp["segment_queue"], inp["pixel_queue"]), 1) if bank else None
    loss = P.pixel_contrast_loss(embed, inp["target"], predict, temperature=cfg["T"], base_temperature=cfg["bT"],
                                 max_samples=cfg["max_samples"], max_views=cfg["max_views"], queue=queue,
                                 per_pair_gather=True)
    loss.backward()
    if bank:
        P.dequeue_and_enqueue(inp["embed"], inp["target"], inp["segment_queue"], inp["segment_queue_ptr"],
                              inp["pixel_queue"], inp["pixel_queue_ptr"], network_stride=cfg["net_stride"],
                              memory_size=cfg["M"], pixel_update_freq=cfg["F"])
    return float(loss.item())


def calibrate_cpu(cfg, bank):
    """Pick the host thread count that runs the port fastest (torch CPU ops stop scaling well before 128 threads)
    and time one single-image step with it."""
    cores = os.cpu_count() or 1
    c1 = dict(cfg); c1["B"] = 1
    inp = make_inputs(c1, 304, None, bank)
    best = (None, float("inf"))
    for th in sorted({min(cores, t) for t in (8, 16, 32, 64, cores)}):
        torch.set_num_threads(th)
        cpu_port_step(inp, c1, bank)
        t0 = time.perf_counter()
        cpu_port_step(inp, c1, bank)
        dt = time.perf_counter() - t0
        if dt < best[1]:
            best = (th, dt)
    torch.set_num_threads(best[0])
    return best


def run_reference(args, cfg, bank, rank, budget_s=150.0):
    if rank != 0:
        return None
    threads, t1 = calibrate_cpu(cfg, bank)
    total = args.steps + args.warmup
    # the reference's backward fills one (B,HW,D) tensor per (image,class) pair: step time grows ~B^2
    B_ref = 1
    for b in (cfg["B"], cfg["B"] // 2, cfg["B"] // 4):
        if b >= 1 and total * t1 * b * b <= budget_s:
            B_ref = b
            break
    c = dict(cfg); c["B"] = B_ref
    inp = make_inputs(c, 304, None, bank)
    for _ in range(args.warmup):
        cpu_port_step(inp, c, bank)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_step(inp, c, bank)
    dt = time.perf_counter() - t0
    ms = dt / args.steps * 1e3
    val = B_ref * args.steps / dt
    sample = (f"{args.steps} steps of the workload geometry at batch {B_ref} (full batch {cfg['B']}; bounded so the run "
              f"fits ~{budget_s:.0f} s) on {threads} of {os.cpu_count()} host threads (fastest of 8/16/32/64/all), fp32 torch "
              "CPU, per-(image,class) gather + autograd backward like the reference")
    return {"impl": "reference", "metric": "contrast-loss fwd+bwd throughput", "value": val, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, bank, B_ref),
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def workload_config(cfg, bank, B=None):
    name = ("HRNet-W48 pixel-contrast + pixel/region memory bank" if bank else "HRNet-W48 pixel-contrast (no memory bank)")
    data = "synthetic Cityscapes 1024x512 19-class"
    if cfg["K"] == 171:
        name, data = "ResNet-101 DeepLab-V3 pixel-contrast + region memory", "synthetic COCO-Stuff 520x520 171-class"
    return {"workload": f"{name}, {data} -> embed {cfg['D']}x{cfg['h']}x{cfg['w']}, "
                        f"batch {B or cfg['B']} per GPU, max_samples {cfg['max_samples']}, max_views {cfg['max_views']}",
            "per_gpu_batch": B or cfg["B"], "parallelism": "dp (images sharded, no data-path collective)" if not bank
            else "dp + one NCCL all_gather of the bank enqueue packet per step",
            "l2": "inputs larger than L2 (embed 268 MB + grad 268 MB per step vs 126 MB L2)" if (B or cfg["B"]) >= 4
            else "L2 flushed between iterations"}


# ---------------------------------------------------------------------------------------------------
# engine arm
# ---------------------------------------------------------------------------------------------------
def stage_timings(cfg, inp, dev, precision, iters=20):
    """CUDA-event time of every C-ABI stage on the launching stream (dominant-kernel roofline)."""
    import ctypes as C
    from contrastiveseg_b200 import _abi, functional as Fn
    lib = _abi.load()
    geom = _abi.Geom(cfg["B"], cfg["D"], cfg["h"], cfg["w"], inp["target"].shape[1], inp["target"].shape[2], cfg["K"],
                     cfg["max_samples"], cfg["max_views"], -1)
    ws = Fn.ContrastWorkspace(dev, geom, 0, 0, 0, 0)
    d = ws.desc
    d.embed, d.labels, d.seg = inp["embed"].data_ptr(), inp["target"].data_ptr(), inp["seg"].data_ptr()
    d.temperature, d.base_temperature, d.seed = cfg["T"], cfg["bT"], 12345
    grad = torch.empty_like(inp["embed"])
    d.grad_embed = grad.data_ptr()
    stream = torch.cuda.current_stream(dev).cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sw = _abi.SweepDesc()
    ms_ = cfg["max_samples"]
    sw.anchors, sw.anchor_cls = ws.anchors_f32.data_ptr(), ws.anchor_meta.data_ptr() + 8 * ms_
    sw.diag_col, sw.plan = ws.anchor_meta.data_ptr() + 12 * ms_, ws.plan.data_ptr()
    sw.a_rows, sw.D, sw.mode = ms_, cfg["D"], 0
    sw.temperature, sw.base_temperature = cfg["T"], cfg["bT"]
    td = _abi.TcDesc()
    td.anchors_bf16, td.anchor_cls = ws.anchors_bf16.data_ptr(), sw.anchor_cls
    td.diag_col, td.plan, td.a_rows, td.D, td.mode = sw.diag_col, sw.plan, ms_, cfg["D"], 0
    td.contrast_norm_bound, td.temperature, td.base_temperature = 1.0, cfg["T"], cfg["bT"]
    g = C.byref(geom)
    if precision == "bf16":
        fwd = lambda: lib.pcl_infonce_tc_fwd(C.byref(td), d.row_m2, d.partials, d.rowstats, d.loss, stream)
        bwd = lambda: lib.pcl_infonce_tc_bwd(C.byref(td), d.row_m2, d.rowstats, None, d.dpartials, d.dA, stream)
    else:
        fwd = lambda: lib.pcl_infonce_fwd(C.byref(sw), d.partials, d.rowstats, d.loss, stream)
        bwd = lambda: lib.pcl_infonce_bwd(C.byref(sw), d.rowstats, None, d.dpartials, d.dA, stream)
    stages = {
        "class_stats": lambda: lib.pcl_class_stats(g, d.labels, d.seg, None, d.keys, d.chunk_pref, d.counts, stream),
        "plan_anchors": lambda: lib.pcl_plan_anchors(g, d.counts, d.plan, stream),
        "select_gather": lambda: lib.pcl_select_gather(g, d.embed, d.keys, d.chunk_pref, d.plan, None, 12345, 0,
                                                       d.anchor_meta, d.anchors_f32, d.anchors_bf16, d.inv_norm,
                                                       d.norm_max, stream),
        "infonce_fwd": fwd,
        "infonce_bwd": bwd,
        "scatter_grad": lambda: lib.pcl_scatter_grad(g, d.plan, d.anchor_meta, d.dA, d.anchors_f32, d.inv_norm, 0,
                                                     d.grad_embed, stream),
    }
    out = {}
    acc = {k: 0.0 for k in stages}
    for it in range(iters + 2):
        flush.zero_()
        for name, fn in stages.items():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            st = fn()
            e1.record()
            _abi.check(st, name)
            torch.cuda.synchronize(dev)
            if it >= 2:
                acc[name] += e0.elapsed_time(e1)
    for k in acc:
        out[k] = acc[k] / iters
    A = int(ws.plan[2].item())
    return out, A


def train_iter_bench(cfg, inp, dev, precision, iters=10):
    """BASELINE.json's second figure, "train iters/sec": one synthetic training iteration = producer fwd -> ContrastCELoss
    (fused seg CE + pixel contrast) -> backward -> SGD step.  The reference's HRNet-W48 cannot travel to the GPU box and
    is outside the hot-path scope, so the producer is a stand-in with the same output contract and tensor sizes
    (lib/models/nets/hrnet.py:76-95: 720-channel stride-4 features -> 19-class head + 720->720->256 projection head
    with the engine's L2-normalise): the number shows the loss step's share of an iteration, not HRNet's speed."""
    import torch.nn as nn
    import contrastiveseg_b200 as cs

    class StandIn(nn.Module):
        def __init__(self, K, D, C=720):
            super().__init__()
            self.body = nn.Sequential(nn.Conv2d(3, 64, 3, 2, 1), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
                                      nn.Conv2d(64, C, 3, 2, 1), nn.BatchNorm2d(C), nn.ReLU(inplace=True))
            self.cls = nn.Conv2d(C, K, 1)
            self.proj = cs.ProjectionHead(C, D, proj="convmlp", bn_type="torchbn")

        def forward(self, x):
            f = self.body(x)
            return {"seg": self.cls(f), "embed": self.proj(f)}

    torch.manual_seed(304)
    net = StandIn(cfg["K"], cfg["D"]).to(dev)
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-4)
    crit = cs.ContrastCELoss(engine_configer(cfg, False, precision)).to(dev)
    H, W = inp["target"].shape[1], inp["target"].shape[2]
    x = torch.randn(cfg["B"], 3, H, W, device=dev)

    def it():
        out = net(x)
        loss = crit(out, inp["target"], with_embed=True)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss
    for _ in range(3):
        it()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(iters):
        loss = it()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / iters
    ok = bool(torch.isfinite(loss).item())
    del net, opt, x
    torch.cuda.empty_cache()
    return {"iters_per_s": 1e3 / ms, "ms_per_iter": ms, "global_batch": cfg["B"], "finite_loss": ok,
            "producer": "stand-in conv stem -> 720ch stride-4 features -> 19-class head + 720-720-256 projection head "
                        "(engine L2-normalise); fp32, SGD momentum 0.9; not HRNet-W48",
            "loss": "ContrastCELoss (fused seg CE + pixel contrast, with_embed=True)"}


# ---------------------------------------------------------------------------------------------------
# CUDA-graph arm: the same step as ONE graph replay (GraphedContrastStep), measured in a child process per rank so
# that nothing it does can touch the process that produced the eager numbers.  The child first proves on the hardware
# that a replay computes exactly what the eager autograd step computes (anchors, loss, dense gradient), then waits for
# the parent's "GO" (sent after a barrier over all ranks) and times K replays with CUDA events.
# ---------------------------------------------------------------------------------------------------
def run_graph_child(args, cfg):
    import contrastiveseg_b200 as cs
    from contrastiveseg_b200 import functional as Fn
    out = {"ok": False}
    try:
        dev = torch.device("cuda:0")                  # the parent narrowed CUDA_VISIBLE_DEVICES to this rank's GPU
        torch.cuda.set_device(dev)
        inp_h = make_inputs(cfg, 304 + args.child_rank, None, False)
        inp = {k: v.to(dev) for k, v in inp_h.items()}
        crit = cs.PixelContrastLoss(engine_configer(cfg, False, args.precision))
        opts = crit.options()
        opts.num_classes = cfg["K"]
        # ---- 1. replay == eager step, on this GPU, before anything is timed ----
        probe = cs.GraphedContrastStep(inp["embed"].clone(), inp["target"], seg=inp["seg"], options=opts)
        for r in range(2):
            loss_g, grad_g = probe.replay()
            torch.cuda.synchronize(dev)
            meta_g, loss_g, grad_g = probe.ws.anchor_meta.clone(), loss_g.clone(), grad_g.clone()
            Fn._step_counter[0] = r
            e = inp["embed"].clone().requires_grad_(True)
            loss_e = cs.pixel_contrast_loss(e, inp["target"], seg=inp["seg"], options=opts)
            ws = Fn.last_workspace(dev)
            loss_e.backward()
            torch.cuda.synchronize(dev)
            same = (torch.equal(ws.anchor_meta, meta_g) and torch.equal(loss_e.detach(), loss_g) and
                    torch.allclose(e.grad, grad_g, rtol=3e-6, atol=0))
            if not same:
                out["why"] = f"replay {r} differs from the eager step"
                print(json.dumps(out), flush=True)
                return
        del probe, e
        torch.cuda.empty_cache()
        # ---- 2. the timed object: static input buffers, one replay per step ----
        step = cs.GraphedContrastStep(inp["embed"], inp["target"], seg=inp["seg"], options=opts)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if cfg["B"] < 4 else None
        for _ in range(max(args.warmup, 3)):
            step.replay()
        torch.cuda.synchronize(dev)
        print("READY", flush=True)
        if sys.stdin.readline().strip() != "GO":
            return
        sampler = ClockSampler(args.child_gpu) if args.child_rank == 0 else None      # nvidia-smi takes the physical index / uuid
        if sampler:
            sampler.start()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        ev[0].record()
        for _ in range(args.steps):
            if flush is not None:
                flush.zero_()
            step.replay()
        ev[1].record()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize(dev)
        while time.perf_counter() - t0 < 0.9:            # keep the GPU under the same load for the clock sampler
            for _ in range(50):
                step.replay()
            torch.cuda.synchronize(dev)
        clocks = None
        if sampler:
            sampler.stop()
            clocks = sampler.summary()
        ms = ev[0].elapsed_time(ev[1]) / args.steps
        # ---- 3. end to end: pinned host inputs copied in every step, loss read back ----
        pin = {k: inp_h[k].pin_memory() for k in ("embed", "seg", "target")}
        e_steps = max(3, min(args.steps, 50))

        def e2e_step():
            for k in pin:
                inp[k].copy_(pin[k], non_blocking=True)
            loss, _ = step.replay()
            return loss.item()
        for _ in range(3):
            e2e_step()
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        torch.cuda.synchronize(dev)
        out.update(ok=True, ms_per_step=ms, host_enqueue_ms_per_step=t_host / args.steps * 1e3,
                   e2e_s_per_step=(time.perf_counter() - t1) / e_steps, e2e_steps=e_steps, clocks=clocks,
                   finite=bool(torch.isfinite(step.loss).item()))
    except Exception as exc:                             # noqa: BLE001
        out["why"] = f"{type(exc).__name__}: {exc}"[:300]
    print(json.dumps(out), flush=True)


class GraphArm:
    """Parent side of the CUDA-graph arm: one child per rank on that rank's GPU (see run_graph_child)."""

    def __init__(self, args, rank, local):
        import subprocess
        env = dict(os.environ)
        vis = [v.strip() for v in env.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
        env["CUDA_VISIBLE_DEVICES"] = vis[local] if local < len(vis) else str(local)
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID"):
            env.pop(k, None)
        self.gpu = env["CUDA_VISIBLE_DEVICES"]
        cmd = [sys.executable, os.path.abspath(__file__), "--graph-child", "--child-rank", str(rank), "--child-gpu", self.gpu,
               "--steps", str(args.steps),
               "--warmup", str(args.warmup), "--precision", args.precision, "--workload", args.workload, "--scaling", args.scaling,
               "--child-world", str(int(os.environ.get("WORLD_SIZE", "1")))]
        import tempfile
        self._err = tempfile.TemporaryFile()             # the child's stderr, quoted in `why` when it dies without an answer
        self.proc = subprocess.Popen(cmd, env=env, stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=self._err)
        self.result = None
        self._buf = b""

    def _stderr_tail(self, n=240):
        try:
            self._err.seek(0)
            txt = self._err.read().decode("utf-8", "replace").strip().replace("\n", " | ")
            return txt[-n:]
        except Exception:                                # noqa: BLE001
            return ""

    def _readline(self, timeout_s):
        """Next protocol line of the child (READY or one JSON object); anything else a library printed is skipped.
        Raw, unbuffered reads: select() must see exactly what has not been consumed yet."""
        import select
        deadline = time.perf_counter() + timeout_s
        fd = self.proc.stdout.fileno()
        while True:
            while b"\n" in self._buf:
                line, self._buf = self._buf.split(b"\n", 1)
                line = line.decode("utf-8", "replace").strip()
                if line == "READY" or line.startswith("{"):
                    return line
            left = deadline - time.perf_counter()
            if left <= 0:
                return None
            r, _, _ = select.select([fd], [], [], left)
            if not r:
                return None
            chunk = os.read(fd, 65536)
            if not chunk:
                return None                              # child closed its stdout (exited)
            self._buf += chunk

    def wait_ready(self, timeout_s=420.0):
        line = self._readline(timeout_s)
        if line == "READY":
            return True
        try:
            self.result = json.loads(line) if line else {"ok": False, "why": "no answer from the graph child (timeout or exit)"
                                                         + (": " + self._stderr_tail() if getattr(self, "_err", None) else "")}
        except ValueError:
            self.result = {"ok": False, "why": f"unexpected output from the graph child: {line[:120]!r}"}
        return False

    def go(self, timeout_s=300.0):
        try:
            self.proc.stdin.write(b"GO\n")
            self.proc.stdin.flush()
            line = self._readline(timeout_s)
            self.result = json.loads(line) if line else {"ok": False, "why": "graph child timed out or died"
                                                         + (": " + self._stderr_tail() if getattr(self, "_err", None) else "")}
        except Exception as exc:                         # noqa: BLE001
            self.result = {"ok": False, "why": f"{type(exc).__name__}: {exc}"[:200]}
        return self.result

    def close(self):
        try:
            if self.proc.poll() is None:
                self.proc.stdin.close()
                self.proc.wait(timeout=4)                # a child that has printed its line exits at once
        except Exception:                                # noqa: BLE001
            pass
        if self.proc.poll() is None:
            self.proc.kill()                             # exactly the child this object started
            try:
                self.proc.wait(timeout=10)
            except Exception:                            # noqa: BLE001
                pass


def graph_arm_measure(args, cfg, rank, world, dev, barrier):
    """Run the CUDA-graph arm (one child per rank, see run_graph_child / GraphArm) and reduce its result over the ranks.
    Every rank takes part in every collective below whatever happens to its own child (no rank may wait alone).
    Returns {"ok": True, value, ms_per_step, e2e_value, ...} or {"ok": False, "why": ...}."""
    arm, ready, why = None, False, None
    try:
        arm = GraphArm(args, rank, dev.index)
        ready = arm.wait_ready()
        if not ready:
            why = (arm.result or {}).get("why", "graph child did not get ready")
    except Exception as exc:                         # noqa: BLE001
        why = f"{type(exc).__name__}: {exc}"[:200]
    flag = torch.tensor([1.0 if ready else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
    if float(flag.item()) > 0.5:
        barrier()                                    # all children are warmed up and wait for GO: start them together
        res = arm.go()
        ok = bool(res.get("ok")) and bool(res.get("finite", False)) and (res.get("ms_per_step") or 0.0) > 1e-3 \
            and (res.get("e2e_s_per_step") or 0.0) > 1e-6
        vals = torch.tensor([1.0 if ok else 0.0, -(res.get("ms_per_step") or 0.0), -(res.get("e2e_s_per_step") or 0.0),
                             -(res.get("host_enqueue_ms_per_step") or 0.0)], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(vals, op=torch.distributed.ReduceOp.MIN)          # MIN of negatives = MAX over ranks
        if float(vals[0].item()) > 0.5:
            g_ms, g_e2e = -float(vals[1].item()), -float(vals[2].item())
            graph = {"ok": True, "ms_per_step": g_ms, "value": world * cfg["B"] / (g_ms / 1e3),
                     "e2e_value": world * cfg["B"] / g_e2e, "e2e_steps": res.get("e2e_steps"),
                     "host_enqueue_ms_per_step": -float(vals[3].item()), "clocks": res.get("clocks"),
                     "check": "on every rank's GPU, before timing: replay == eager autograd step (same anchors, same loss "
                              "bits, gradient within 3e-6 relative: scatter-only backward vs fused writer)"}
        else:
            graph = {"ok": False, "why": res.get("why", "a rank's graph child failed")}
    else:
        graph = {"ok": False, "why": why or "another rank's graph child did not get ready"}
    if arm is not None:
        arm.close()
    return graph


def tensor_sweep_roofline(dev, peaks, A=16384, N=65536, iters=5):
    """The dense contraction alone (similarity + negative-sum sweep on tcgen05) at one S4 point (BASELINE configs[4]):
    algorithmic FLOPs 2*A*N*D over the CUDA-event time, vs the measured bf16 peak."""
    from contrastiveseg_b200 import functional as Fn
    g = torch.Generator().manual_seed(7)
    a = torch.nn.functional.normalize(torch.randn(A, 256, generator=g), dim=1).to(dev)
    c = torch.nn.functional.normalize(torch.randn(N, 256, generator=g), dim=1).to(dev)
    ya = torch.sort(torch.randint(0, 19, (A,), generator=g)).values.to(dev)
    yc = torch.sort(torch.randint(0, 19, (N,), generator=g)).values.to(dev)
    c16 = Fn.to_bf16_rows(c, -(-N // 256) * 256)
    diag = torch.arange(A, device=dev) % N
    out = {}
    for name, neg_only in (("neg_sweep", True), ("forward", False)):
        run = lambda: Fn.infonce_tc_forward(a, ya, contrast_bf16=c16, contrast_cls=yc, n_cols=N, diag_col=diag,
                                            temperature=0.07, base_temperature=0.07, neg_only=neg_only)
        for _ in range(3):
            res = run()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); res = run(); e1.record()
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        out[name] = ts[len(ts) // 2]
    state = res
    loss, st, stt = state
    tb = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); Fn.infonce_tc_backward(stt, st); e1.record()
        torch.cuda.synchronize(dev)
        tb.append(e0.elapsed_time(e1))
    tb.sort()
    out["backward"] = tb[len(tb) // 2]
    fl = 2.0 * A * N * 256
    ach = fl / (out["neg_sweep"] * 1e-3) / 1e12
    return {"kernel": "k_tc_fwd<NEG> (anchor x bank similarity + negative-sum sweep, tcgen05)", "bound": "tensor",
            "workload": f"S4 sweep point A={A} x N={N}, D=256, bf16 operands", "achieved": ach,
            "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
            "peak_source": peaks["source"], "traffic": None, "ms": out,
            "forward_tflops": fl / (out["forward"] * 1e-3) / 1e12, "backward_tflops": fl / (out["backward"] * 1e-3) / 1e12}


def run_engine(args, cfg, bank, rank, world, dev):
    import contrastiveseg_b200 as cs
    from contrastiveseg_b200 import _abi
    _abi.load()
    peaks = load_peaks()
    inp_h = make_inputs(cfg, 304 + rank, None, bank)
    inp = {k: v.to(dev) for k, v in inp_h.items()}
    cfgr = engine_configer(cfg, bank, args.precision)
    crit = cs.PixelContrastLoss(cfgr)
    mbank = None
    if bank:
        mbank = cs.MemoryBank(cfg["K"], cfg["M"], cfg["D"], with_shadow=True).to(dev)
        mbank.segment_queue.copy_(inp["segment_queue"]); mbank.pixel_queue.copy_(inp["pixel_queue"])
        mbank.sync_shadow()
    embed = inp["embed"].clone().requires_grad_(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if cfg["B"] < 4 else None

    graphed = {}

    def step(e, tgt, seg):
        if args.graph:
            # one CUDA-graph replay per step (GraphedContrastStep: static tensors, device-side sampling counter)
            key = (e.data_ptr(), tgt.data_ptr(), seg.data_ptr())
            g = graphed.get(key)
            if g is None:
                kw = dict(segment_queue=mbank.segment_queue, pixel_queue=mbank.pixel_queue,
                          bank_shadow=mbank.shadow) if bank else {}
                opts = crit.options()
                opts.num_classes = cfg["K"]
                g = graphed[key] = cs.GraphedContrastStep(e.detach(), tgt, seg=seg, options=opts, **kw)
            loss, _ = g.replay()
            if bank:
                mbank.enqueue(e.detach(), tgt, network_stride=cfg["net_stride"], pixel_update_freq=cfg["F"])
            return loss
        e.grad = None
        queue = (mbank.segment_queue, mbank.pixel_queue) if bank else None
        loss = crit(e, tgt, seg=seg, queue=queue, bank_shadow=mbank.shadow if bank else None)
        if bank:
            mbank.enqueue(e.detach(), tgt, network_stride=cfg["net_stride"], pixel_update_freq=cfg["F"])
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step(embed, inp["target"], inp["seg"])
    # one nvidia-smi poller for the whole job (rank 0, its own GPU): every poll takes driver locks that stall kernel
    # launches of ALL ranks, so N pollers would perturb a host-launch-bound step
    sampler = ClockSampler(dev.index)
    barrier()
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    t_wall0 = time.perf_counter()
    ev[0].record()
    for _ in range(args.steps):
        if flush is not None:
            flush.zero_()
        step(embed, inp["target"], inp["seg"])
    ev[1].record()
    t_enqueue = time.perf_counter() - t_wall0        # host time to enqueue the K steps (no sync inside)
    barrier()
    # nvidia-smi samples every 200 ms; a short timed region (K steps of ~0.2 ms) would see no sample, so the SAME
    # step loop keeps running (untimed) until the sampler has covered >= 0.9 s under load
    clock_window = "timed region"
    while time.perf_counter() - t_wall0 < 0.9:
        clock_window = "timed region + untimed continuation of the same step loop to 0.9 s"
        for _ in range(50):
            step(embed, inp["target"], inp["seg"])
        torch.cuda.synchronize(dev)
    sampler.stop()
    t_ms = ev[0].elapsed_time(ev[1])
    tt = torch.tensor([t_ms], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        # diagnostics: every rank's device time and host enqueue time per step (which rank sets the max, and why)
        mine = torch.tensor([t_ms / args.steps, t_enqueue / args.steps * 1e3], dtype=torch.float64, device=dev)
        allv = torch.empty(2 * world, dtype=torch.float64, device=dev)
        torch.distributed.all_gather_into_tensor(allv, mine)
        per_rank = {"ms_per_step": [round(v, 5) for v in allv.view(world, 2)[:, 0].tolist()],
                    "host_enqueue_ms_per_step": [round(v, 5) for v in allv.view(world, 2)[:, 1].tolist()]}
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
    t_ms = float(tt.item())
    value = world * cfg["B"] * args.steps / (t_ms / 1e3)

    # ---- e2e: host (pinned) buffers in, loss scalar out, every step ----
    pin = {k: inp_h[k].pin_memory() for k in ("embed", "seg", "target")}
    dbuf = {k: torch.empty_like(inp[k]) for k in ("embed", "seg", "target")}
    h2d = sum(pin[k].numel() * pin[k].element_size() for k in pin)

    def e2e_step():
        for k in pin:
            dbuf[k].copy_(pin[k], non_blocking=True)
        e = dbuf["embed"].requires_grad_(True)
        loss = step(e, dbuf["target"], dbuf["seg"])
        dbuf["embed"] = e.detach()
        return loss.item()                          # D2H read of the step's result
    e_steps = max(3, min(args.steps, 50))
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        e2e_step()
    barrier()
    e_dt = time.perf_counter() - t0
    et = torch.tensor([e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(et, op=torch.distributed.ReduceOp.MAX)
    e2e_val = world * cfg["B"] * e_steps / float(et.item())

    # ---- CUDA-graph arm (child process per rank; the eager numbers above are already final) ----
    graph = None
    if not bank and not args.graph and not args.no_graph_arm and not os.environ.get("PCL_BENCH_NO_GRAPH_ARM"):
        graph = graph_arm_measure(args, cfg, rank, world, dev, barrier)
    if rank != 0:
        return None
    # ---- per-stage times + roofline of the dominant stage (rank 0, N-independent) ----
    roof = None
    stages = {}
    if not bank:
        stages, A = stage_timings(cfg, inp, dev, args.precision)
        # roofline of the one single-pass bandwidth-bound kernel of the step, the dense-gradient writer (the other
        # stages are chains of short latency-bound launches at A <= 1024; stage_ms lists them all)
        dom = "scatter_grad"
        BDHW4 = cfg["B"] * cfg["D"] * cfg["h"] * cfg["w"] * 4
        alg_bytes = {
            "scatter_grad": BDHW4 + 2 * A * cfg["D"] * 4,
            "class_stats": cfg["B"] * cfg["h"] * cfg["w"] * (cfg["K"] * 4 + 8 + 2),
            "select_gather": A * cfg["D"] * (4 + 4 + 2) + cfg["B"] * cfg["h"] * cfg["w"] * 2,
        }
        if dom in alg_bytes:
            ach = alg_bytes[dom] / (stages[dom] * 1e-3) / 1e9
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"], "traffic": ncu_traffic("k_zero_scatter"),
                    "peak_source": peaks["source"], "algorithmic_bytes": alg_bytes[dom],
                    "kernel_name": "k_zero_scatter" if dom == "scatter_grad" else dom}
        else:
            flops = 2.0 * A * A * cfg["D"] * (2 if dom == "infonce_fwd" else 2)
            ach = flops / (stages[dom] * 1e-3) / 1e12
            roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": ach / peaks["bf16_tflops"], "traffic": None, "peak_source": peaks["source"],
                    "note": "A=N<=1024 is the launch/latency-bound regime of the sweep (SURVEY §8d); tensor_roofline reports the S4 point"}
    # ---- cpu baseline (rank 0, N=1 only): bounded sample of the same workload ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads, t1 = calibrate_cpu(cfg, bank)
        Bc = cfg["B"] if 3 * t1 * cfg["B"] ** 2 <= 40.0 else max(1, cfg["B"] // 2) if 3 * t1 * (cfg["B"] // 2) ** 2 <= 40.0 else 1
        cc = dict(cfg); cc["B"] = Bc
        ih = make_inputs(cc, 304, None, bank) if Bc != cfg["B"] else inp_h
        cpu_port_step(ih, cc, bank)
        n_cpu = 2
        t0 = time.perf_counter()
        for _ in range(n_cpu):
            cpu_port_step(ih, cc, bank)
        cdt = time.perf_counter() - t0
        cpu = {"value": Bc * n_cpu / cdt, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"{n_cpu} steps at batch {Bc} (of {cfg['B']}) after 1 warm-up, fp32 torch CPU on {threads} of "
                         f"{os.cpu_count()} host threads (fastest setting), {cdt / n_cpu * 1e3:.0f} ms/step"}
    tens = tensor_sweep_roofline(dev, peaks) if (world == 1 and cfg["D"] == 256) else None
    # whole ContrastCELoss.forward + backward (seg CE + contrast): the reference's "Loss Time" scope, with the fused
    # up-sample + CE kernels (§8f row 1) and with the PyTorch seg-CE ops
    wrapper = None
    if world == 1 and not bank:
        wrapper = {}
        for name, fused in (("fused_seg_ce_ms", True), ("torch_seg_ce_ms", False)):
            cw = engine_configer(cfg, bank, args.precision)
            cw.add(["contrast", "fused_seg_ce"], fused)
            mod = cs.ContrastCELoss(cw).to(dev)
            seg_l = inp["seg"].clone().requires_grad_(True)

            def wstep():
                embed.grad = None; seg_l.grad = None
                mod({"seg": seg_l, "embed": embed}, inp["target"], with_embed=True).backward()
            for _ in range(5):
                wstep()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            e0.record()
            for _ in range(30):
                wstep()
            e1.record()
            torch.cuda.synchronize(dev)
            wrapper[name] = e0.elapsed_time(e1) / 30
    # "train iters/sec" half of BASELINE.json's metric (N=1, stand-in producer); never allowed to break the bench line
    train = None
    if world == 1 and not bank and not os.environ.get("PCL_BENCH_NO_TRAIN_ITER"):
        try:
            train = train_iter_bench(cfg, inp, dev, args.precision)
            if wrapper:
                train["loss_step_ms"] = wrapper.get("fused_seg_ce_ms")
                train["loss_step_share"] = wrapper.get("fused_seg_ce_ms") / train["ms_per_iter"]
        except Exception as exc:                      # noqa: BLE001
            train = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    launches_per_step = (8 if args.precision == "bf16" else 10) + (4 if bank else 0)   # our kernels per step (memsets not counted)
    if args.graph:
        launches_per_step += 1                       # + the device-side rank draw (pcl_step_ranks); one graph launch per step
    # headline = the faster of the two ways the public API offers to run the step: eager autograd call, or one CUDA-graph
    # replay (GraphedContrastStep) — the latter only if every rank proved replay == eager on its GPU and finished
    eager = {"value": value, "ms_per_step": t_ms / args.steps, "e2e_value": e2e_val,
             "host_enqueue_ms_per_step": t_enqueue / args.steps * 1e3, "kernels_per_step": launches_per_step}
    use_graph = bool(graph and graph.get("ok") and graph["ms_per_step"] < t_ms / args.steps)
    clocks_out = dict(sampler.summary(), window=clock_window)
    if use_graph:
        value, e2e_val = graph["value"], graph["e2e_value"]
        e_steps = graph.get("e2e_steps") or e_steps
        t_ms = graph["ms_per_step"] * args.steps
        launches_per_step += 2                       # + rank draw, + the reduction kernel of the scatter-only backward
        if graph.get("clocks") and graph["clocks"].get("samples"):
            clocks_out = dict(graph["clocks"], window="graph arm: timed region + untimed continuation of the same replay loop to 0.9 s")
    return {"metric": "contrast-loss fwd+bwd throughput", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic", "config": dict(workload_config(cfg, bank), step="one CUDA-graph replay per step (GraphedContrastStep)" if use_graph else "eager autograd call per step"), "clocks": clocks_out,
            "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "steps": e_steps},
            "gpu_launches": launches_per_step * args.steps, "roofline": roof, "cpu_baseline": cpu,
            "stage_ms": stages, "host_enqueue_ms_per_step": graph["host_enqueue_ms_per_step"] if use_graph else t_enqueue / args.steps * 1e3, "tensor_roofline": tens, "contrast_ce_wrapper": wrapper, "per_rank": per_rank, "train_iter": train, "precision": args.precision, "cuda_graph": bool(args.graph) or use_graph,
            "eager": eager, "graph_arm": graph, "impl": "engine"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="s1", choices=["s1", "s2", "s3"],
                    help="s1 = BASELINE configs[1] (headline), s2 = configs[2] (memory bank), s3 = configs[3] (171 classes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): the workload's batch per GPU; strong: the workload's batch split over the ranks "
                         "(global batch fixed, as lib/datasets/data_loader.py:137 does), SURVEY §8d asks for both")
    ap.add_argument("--no-graph-arm", action="store_true", help="skip the CUDA-graph arm (child process per rank)")
    ap.add_argument("--graph-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--child-rank", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--child-world", type=int, default=1, help=argparse.SUPPRESS)
    ap.add_argument("--child-gpu", default="0", help=argparse.SUPPRESS)
    ap.add_argument("--graph", action="store_true",
                    help="run the step as one CUDA-graph replay (GraphedContrastStep) instead of the eager autograd call")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="InfoNCE sweeps: bf16 operands on tcgen05 tensor cores (default) or the exact fp32 SIMT sweep")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")) if not args.graph_child else args.child_world
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = dict({"s1": S1, "s2": S2, "s3": S3}[args.workload])
    bank = args.workload in ("s2", "s3")
    if args.scaling == "strong":
        if cfg["B"] % world != 0:
            raise SystemExit(f"--scaling strong: global batch {cfg['B']} is not divisible by {world} ranks")
        cfg["B"] //= world
    if os.environ.get("PCL_BENCH_TINY"):            # contract tests on small hosts: same code path, toy geometry
        cfg.update(B=2, D=32, h=16, w=16, K=5, stride=2, block=8, max_samples=32, max_views=4)
        if bank:
            cfg.update(M=8, F=2, net_stride=2)
    if args.graph_child:
        run_graph_child(args, cfg)
        return
    if args.impl == "reference":
        if args.steps == 200 and args.warmup == 10:      # defaults are sized for the GPU arm
            args.steps, args.warmup = 3, 1
        res = run_reference(args, cfg, bank, rank)
        if res is not None:
            print(json.dumps(res), flush=True)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl engine needs a CUDA device (the engine has no CPU path)")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        pin_to_gpu_local_cpus(dev.index)
        torch.distributed.init_process_group("nccl", device_id=dev)
    res = run_engine(args, cfg, bank, rank, world, dev)
    if res is not None:
        print(json.dumps(res), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
