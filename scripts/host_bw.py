"""host memory write bandwidth with N threads (decides whether a packed device->host witness transfer
with host-side expansion can beat the plain PCIe copy)"""
import sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
import torch

total_gb = 8
for pinned in (False, True):
    buf = torch.empty(total_gb * (1 << 30), dtype=torch.uint8, pin_memory=pinned).numpy()
    buf[:] = 1
    src = np.zeros(64 << 20, dtype=np.uint8)
    for nt in (8, 32, 64, 128):
        chunk = buf.size // nt
        def work(i):
            d = buf[i * chunk:(i + 1) * chunk]
            for o in range(0, chunk, src.size):
                n = min(src.size, chunk - o)
                np.copyto(d[o:o + n], src[:n])
        with ThreadPoolExecutor(nt) as ex:
            t0 = time.time(); list(ex.map(work, range(nt))); dt = time.time() - t0
        print("pinned=%s threads=%d: %.1f GB/s" % (pinned, nt, total_gb * 1.0737 / dt))
