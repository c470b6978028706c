"""Host-clock A/B of the sweep-ordering variants: srl_sweep_set_device (device-to-device copy + ordering) x REP, one sync.
cluster_order: 0 CUB, 1 cluster kernel (registers + CTA-local digit order), 2 registers + direct scatter, 3 first version."""
import sys, time
import numpy as np, torch
from sr_livo_b200 import lio
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 300
L = lio.LioOptimization(max_voxels=1 << 12, sweep_capacity=n)
rng = np.random.default_rng(3)
raw = torch.from_numpy(rng.uniform(-60.0, 60.0, size=(n, 3))).cuda()
torch.cuda.synchronize()
for mode in (0, 1, 2, 3, 0, 1):
    L.ctx.set_option("cluster_order", mode)
    for _ in range(8):
        L.sweep.set_device(raw.data_ptr(), n)
    L.ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(rep):
        L.sweep.set_device(raw.data_ptr(), n)
    L.ctx.synchronize()
    dt = (time.perf_counter() - t0) / rep * 1e6
    print("cluster_order=%d  n=%d  %.1f us per set_device+order  (impl counter %d)" % (mode, n, dt, L.ctx.counter("cluster_order_active")), flush=True)
