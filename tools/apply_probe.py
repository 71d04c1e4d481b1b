"""Time of pcl_bank_apply for the packets of W ranks on ONE GPU (BASELINE configs[2] geometry per rank: one image, K = 19,
M = 5000, F = 10, D = 256).  The packets are W copies of one real packet (the kernel's work does not depend on whose
rows they are).  Shows that the apply no longer grows with the number of ranks (last-writer-wins, csrc/pcl_bank.cu)."""
import ctypes as C
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import contrastiveseg_b200 as cs
from contrastiveseg_b200 import _abi, bank as bank_mod, functional as Fn
from contrastiveseg_b200.synth import make_bank, make_contrast_batch

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
lib = _abi.load()
K, D, M, F, stride = 19, 256, 5000, 10, 4
data = make_contrast_batch(B=1, D=D, h=128, w=256, num_classes=K, img_stride=4, block=32, seed=3)
b0 = make_bank(K, M, D, 1)
names = ("segment_queue", "segment_queue_ptr", "pixel_queue", "pixel_queue_ptr")
bufs = [b0[k].clone().to(dev) for k in names]
keep = []
orig = bank_mod.gather_packets
bank_mod.gather_packets = lambda pk, group=None, out=None: (keep.append(pk.clone()), pk.view(1, -1))[1]
cs.dequeue_and_enqueue(data["embed"].to(dev), data["target"].to(dev), *bufs, network_stride=stride, memory_size=M,
                       pixel_update_freq=F)
bank_mod.gather_packets = orig
packet = keep[0].view(-1)
g = _abi.BankGeom(1, D, 128, 256, 512, 1024, K, M, stride, F)
shadow = torch.zeros((bank_mod.shadow_rows(K, M), D), dtype=torch.bfloat16, device=dev)
for world in (1, 2, 8, 32, 64):
    packets = packet.repeat(world, 1).contiguous()
    fn = lambda: _abi.check(lib.pcl_bank_apply(C.byref(g), packets.data_ptr(), world, bufs[0].data_ptr(),
                                               bufs[1].data_ptr(), bufs[2].data_ptr(), bufs[3].data_ptr(),
                                               shadow.data_ptr(), Fn._stream_ptr(dev)), "pcl_bank_apply")
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"world": world, "apply_us": e0.elapsed_time(e1) / 200 * 1e3, "packet_kb": packet.numel() * 4 / 1024}),
          flush=True)
