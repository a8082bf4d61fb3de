"""Is the training step GPU-bound?  Host enqueue time per step vs device time, and the same forward+backward
replayed as a CUDA graph (no launch gaps).  python tools/step_gaps.py [cfg3|cfg2|cfg3p]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
w = bench.WORKLOADS[wl]
dev = torch.device("cuda", 0)
model, opt, step_fn = bench.build_model(w, dev)
x, y = bench.make_batch(w, w["batch"], 1, dev)


def step():
    opt.zero_grad(set_to_none=True)
    loss = step_fn(x, y)
    loss.backward()
    opt.step()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
N = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(N):
    step()
e1.record()
t_enq = (time.perf_counter() - t0) / N * 1e3
torch.cuda.synchronize()
print(f"{wl}: device {e0.elapsed_time(e1) / N:.3f} ms/step, host enqueue {t_enq:.3f} ms/step")

# forward + backward as one CUDA graph (optimizer outside)
try:
    params = [p for p in model.parameters()]
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            opt.zero_grad(set_to_none=True)
            step_fn(x, y).backward()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    opt.zero_grad(set_to_none=True)
    with torch.cuda.graph(g):
        loss = step_fn(x, y)
        loss.backward()
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
        opt.step()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(N):
        g.replay()
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    print(f"{wl}: graph replay (fwd+bwd) + optimizer {e0.elapsed_time(e1) / N:.3f} ms/step, loss {loss.item():.5f}")
    e0.record()
    for _ in range(N):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{wl}: graph replay only {e0.elapsed_time(e1) / N:.3f} ms/step")
except Exception as e:  # noqa: BLE001
    print("graph capture failed:", type(e).__name__, str(e)[:300])
