"""Pipeline timeline of CTA 0 of the cp.async tcgen05 forward kernel (development aid)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TODA_TC_FEED"] = "cpasync"
import numpy as np, torch
from toda_b200 import ops, _C
from tests import parity_utils as PU
cin, cout, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
shape, batch = ([int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [21, 400, 400]), 2
feats, idx = PU.random_sparse(3, batch, shape, n, cin)
x = torch.from_numpy(feats).cuda()
index = ops.OccupancyIndex(batch, shape, torch.device("cuda", 0), "dbg")
index.insert(torch.from_numpy(idx).cuda()); index.build(n)
rb = ops.rulebook_subm(index, [3, 3, 3])
w = torch.randn(cout, 3, 3, 3, cin, device="cuda") * 0.1
L = _C.lib()
buf = torch.zeros(8 * 256, dtype=torch.int64, device="cuda")
for _ in range(3): ops.sparse_conv(x, w, None, rb, ops.CONV_BF16)
L.toda_debug_set_timeline.argtypes = [ctypes.c_void_p]
L.toda_debug_set_timeline(ctypes.c_void_p(buf.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.sparse_conv(x, w, None, rb, ops.CONV_BF16); e1.record(); torch.cuda.synchronize()
L.toda_debug_set_timeline(None)
t = buf.cpu().numpy().reshape(8, 256).astype(np.int64)[:7]
t0 = t[4, 1]
names = ["M:fenced", "M:mma0", "M:mma_last", "P:signalled", "M:wait full", "M:full ok", "M:committed"]
print("conv %d->%d n=%d: %.3f ms (incl. pre-passes)" % (cin, cout, n, e0.elapsed_time(e1)))
print("chunk " + " ".join("%12s" % s for s in names))
for g in range(8, 40):
    print("%5d " % g + " ".join("%12d" % (t[r, g] - t0) for r in range(7)))
sig = t[3, 8:250]
print("producer chunk period (cycles): median %d  mean %d;  MMA wait (full ok - wait full) median %d  mean %d;  pairs/row %.2f" % (
    np.median(np.diff(sig)), np.mean(np.diff(sig)), np.median(t[5, 8:250] - t[4, 8:250]), np.mean(t[5, 8:250] - t[4, 8:250]),
    float((rb.nbr_fwd >= 0).sum().item()) / n))
