"""Per-call latency of the reference call shape (one x[N] -> [K]) through the drop-in API.  GPU only."""
import time, numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from qkan_implementation_b200 import QKANLayer
from oracle import qkan_oracle as o
np.random.seed(42)
x = np.random.uniform(-1, 1, 4); W = [np.random.uniform(-1, 1, 16) for _ in range(4)]
layer = QKANLayer(4, 4, 3)
for _ in range(20): y = layer.forward(x, W)
t0 = time.perf_counter()
for _ in range(2000): y = layer.forward(x, W)
dt = (time.perf_counter() - t0) / 2000
print(f"single-sample forward (reference call shape, numpy in/out): {dt*1e6:.1f} us per call")
Wa = np.array(W)
t0 = time.perf_counter()
for _ in range(2000): r = o.forward_reference_style(x, W, 4, 4, 3)
print(f"oracle reference-style per call: {(time.perf_counter()-t0)/2000*1e6:.1f} us")
xd = torch.from_numpy(x).cuda(); Wd = torch.from_numpy(Wa).cuda()
for _ in range(20): layer.forward(xd, Wd)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(2000): layer.forward(xd, Wd)
torch.cuda.synchronize()
print(f"device tensors, async: {(time.perf_counter()-t0)/2000*1e6:.1f} us per call")
