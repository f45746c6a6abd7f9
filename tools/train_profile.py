"""One training micro-step (BASELINE.json configs[4]: micro-batch 4, 640x512) for profilers: warm-up steps outside the
capture range, then ONE step between cudaProfilerStart / Stop (ncu --profile-from-start off). Also prints the per-symbol
CUDA-event time shares of the library's own launches (imagdressing_b200._lib.profile_launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from imagdressing_b200 import _lib, train  # noqa: E402

B = int(os.environ.get("B", "4"))
H, W = int(os.environ.get("H", "640")), int(os.environ.get("W", "512"))
dev = torch.device("cuda:0")
sd, opt, sched = bench.build_train(dev)
x = bench.synth_train_batch(B, dev, 0, False, H // 8, W // 8)
for _ in range(2):
    train.train_step(sd, sched, optimizer=opt, **x)
torch.cuda.synchronize()
if os.environ.get("EVENTS", "1") == "1":
    with _lib.profile_launches() as rec:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        train.train_step(sd, sched, optimizer=opt, **x)
        e1.record()
    by = rec.by_key()
    total = sum(ms for _, ms in by.values())
    print(f"step (eager, event-bracketed) {e0.elapsed_time(e1):.1f} ms; library launches {sum(n for n, _ in by.values())}, "
          f"their kernel time {total:.1f} ms")
    fam = {}
    for k, (n, ms) in by.items():
        f = k.split("(")[0]
        a = fam.setdefault(f, [0, 0.0])
        a[0] += n
        a[1] += ms
    for f, (n, ms) in sorted(fam.items(), key=lambda t: -t[1][1]):
        print(f"  {f:36s} {n:5d} launches {ms:9.2f} ms  {100 * ms / total:5.1f} %")
    print("  top launch keys:")
    for k, (n, ms) in sorted(by.items(), key=lambda t: -t[1][1])[:25]:
        print(f"    {ms:8.2f} ms {n:4d} x {k}")
torch.cuda.synchronize()
torch.cuda.profiler.start()
train.train_step(sd, sched, optimizer=opt, **x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
