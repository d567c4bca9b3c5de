import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from dyn_res_pile_manip_b200 import ops, synthetic
g = synthetic.make_goal("bar")
coords = np.argwhere(g < 0.5)[:, ::-1].astype(np.float32)
pts = torch.as_tensor(np.ascontiguousarray(coords), device="cuda")
for count in (500, 1500):
    ops.fps(pts, count, 0); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        picked, idx, rad = ops.fps(pts, count, 0)
    torch.cuda.synchronize()
    ref, r = synthetic.fps_np(coords, count, 0)
    print("n", len(coords), "count", count, "ms", (time.perf_counter() - t) / 3 * 1e3, "match", np.array_equal(picked.cpu().numpy(), ref), float(rad), float(r))
# per-pick latency as a function of the cloud size (one set)
rng = np.random.RandomState(0)
for n in (512, 2048, 8192, 28800):
    p = torch.as_tensor(rng.uniform(0, 700, size=(n, 2)).astype(np.float32), device="cuda")
    ops.fps(p, 256, 0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.fps(p, 256, 0)
    e1.record(); torch.cuda.synchronize()
    print("n", n, "us per pick %.2f" % (e0.elapsed_time(e1) * 1e3 / 256))
