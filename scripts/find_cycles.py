"""Development aid: which reference cycles does one training step of the hot path leave behind (objects only the cyclic
GC can free pin device memory until it runs)?"""
import gc, os, sys
from collections import Counter
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
vfe, net, hc = bench.build_hot_path(dev, "bf16")
pts, offs = bench.host_batches(0, 1)[0]
pts, offs = pts.to(dev), offs.to(dev)


def step():
    bd = vfe({"points": pts, "point_frame_offsets": offs, "batch_size": bench.FRAMES_PER_GPU})
    bd = hc(net(bd))
    sf = bd["spatial_features"]
    (sf * 1e-6).sum().backward()


for _ in range(2):
    step()
gc.collect()
gc.disable()
torch.cuda.synchronize()
m0 = torch.cuda.memory_allocated()
step()
torch.cuda.synchronize()
m1 = torch.cuda.memory_allocated()
gc.set_debug(gc.DEBUG_SAVEALL)
n = gc.collect()
print("memory held after one step with the GC off: %.1f MB; unreachable objects: %d" % ((m1 - m0) / 2**20, n))
print(Counter(type(o).__name__ for o in gc.garbage).most_common(25))
tens = [o for o in gc.garbage if torch.is_tensor(o)]
print("tensors in cycles:", len(tens), "MB:", sum(t.numel() * t.element_size() for t in tens) / 2**20)
for o in gc.garbage:
    if type(o).__name__ not in ("Tensor", "tuple", "dict", "list", "cell", "function", "int", "float"):
        refs = [type(r).__name__ for r in gc.get_referents(o)][:12]
        print(type(o).__module__, type(o).__name__, "->", refs)
