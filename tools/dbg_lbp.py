"""small LBP run for compute-sanitizer: one 640x480 frame, compare with the oracle"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import _libs as L
import grayskull_b200 as g
from grayskull_b200 import api
a = L.natural_like(640, 480, 3)
cas = g.load_cascade()
d = torch.from_numpy(a[None]).cuda()
ii = api.integral_batch(d)
rects, counts = api.lbp_detect_batch(cas, ii, 4096, 1.1, 1.0, 4.0, 2)
torch.cuda.synchronize()
got = api.rects_to_numpy(rects, counts)[0]
O = L.oracle(); hc = L.HostCascade()
iio = np.empty(a.shape, np.uint32); O.gso_integral(L.ptr(a), 640, 480, L.ptr(iio))
r = np.zeros(4096, L.RECT_DTYPE)
n = O.gso_lbp_detect(hc.ptr, L.ptr(iio), 640, 480, L.ptr(r), 4096, 1.1, 1.0, 4.0, 2)
print("gpu", len(got), "oracle", n, "equal", got.tobytes() == r[:n].tobytes())
