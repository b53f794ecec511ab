"""Host-side profile of one pipelined training step (development aid): where do the ~8 ms of Python go?
usage: host_profile.py [workload]"""
import sys, os, cProfile, pstats, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "roundup_power2_divisions:8")
import numpy as np, torch
import bench
bench.WORKLOAD = sys.argv[1] if len(sys.argv) > 1 else "nus_0075"
from toda_b200.pipeline import InputPipeline
from toda_b200.dist import FlatGradBucket
dev = torch.device("cuda", 0)
vfe, net, hc = bench.build_hot_path(dev, "bf16")
bucket = FlatGradBucket(net.parameters())
hosts = bench.host_batches(0, 3, dev)
devs = [(p.to(dev), o.to(dev)) for p, o in hosts]
pipe = InputPipeline(vfe, net, dev)
cot = None
def back(h):
    global cot
    bucket.zero()
    bd = hc(net(pipe.consume(h)))
    sf = bd["spatial_features"]
    if cot is None:
        cot = torch.randn(sf.shape, device=dev) / sf.numel()
    (sf * cot).sum().backward()
def front(i):
    p, o = devs[i % 3]
    return pipe.submit({"points": p, "point_frame_offsets": o, "batch_size": 4}, inputs_pending=False)
h = front(0)
for i in range(8):
    back(h); h = front(i + 1)
torch.cuda.synchronize()
import gc; gc.collect(); gc.freeze()
t_f = t_b = 0.0
N = 20
for i in range(N):
    t0 = time.perf_counter(); back(h); t1 = time.perf_counter(); h = front(i + 1); t2 = time.perf_counter()
    t_b += t1 - t0; t_f += t2 - t1
torch.cuda.synchronize()
print("host ms per step WITHOUT profiler: fwd+bwd %.2f  front %.2f" % (t_b / N * 1e3, t_f / N * 1e3))
t_f = t_b = 0.0
pr = cProfile.Profile()
for i in range(N):
    t0 = time.perf_counter()
    pr.enable(); back(h); pr.disable()
    t1 = time.perf_counter()
    h = front(i + 1)
    t2 = time.perf_counter()
    t_b += t1 - t0; t_f += t2 - t1
torch.cuda.synchronize()
print("host ms per step: fwd+bwd %.2f  front %.2f" % (t_b / N * 1e3, t_f / N * 1e3))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print("\n".join(l[:150] for l in s.getvalue().splitlines()[:60]))
