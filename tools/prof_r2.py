#!/usr/bin/env python
"""Small driver for ncu captures: runs ONE op a few times on inputs larger than L2.
usage: prof_r2.py blur_sobel|fast|orb|lbp|integral|resize|box15 [batch]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import grayskull_b200 as g
from grayskull_b200 import api
g.lib().gs_b200_set_device(0)
op = sys.argv[1]
torch.manual_seed(1)
if op in ("blur_sobel", "box15", "blur15", "integral", "resize", "resize_odd", "blur5", "sobel"):
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    src = torch.randint(0, 256, (n, 4096, 4096), dtype=torch.uint8, device="cuda")
    out = torch.zeros_like(src)
    for _ in range(3):
        if op == "blur_sobel": api.blur_sobel_batch(src, 5, out=out)
        elif op == "blur5": api.blur_batch(src, 5, out=out)
        elif op == "sobel": api.sobel_batch(src, out=out)
        elif op == "blur15": api.blur_batch(src, 15, out=out)
        elif op == "box15": api.adaptive_threshold_batch(src, 15, 5, out=out)
        elif op == "integral": ii = api.integral_batch(src)
        elif op == "resize": api.resize_batch(src, 2048, 2048)
        elif op == "resize_odd": api.resize_batch(src, 2560, 1440)
elif op == "integral_uhd":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 150          # 150 x 4 strips = 600 CTAs: the 1024-column strip kernel
    src = torch.randint(0, 256, (n, 2160, 3840), dtype=torch.uint8, device="cuda")
    ii = torch.empty((n, 2160, 3840), dtype=torch.int32, device="cuda")
    for _ in range(3):
        api.integral_batch(src, out=ii)
elif op in ("fast", "orb"):
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    src = api.blur_batch(torch.randint(0, 256, (n, 1080, 1920), dtype=torch.uint8, device="cuda"), 3)
    sm = torch.zeros_like(src)
    for _ in range(3):
        api.orb_extract_batch(src, 1250, 20, scoremap=sm)
elif op == "lbp":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    cas = g.load_cascade()
    src = api.blur_batch(torch.randint(0, 256, (n, 2160, 3840), dtype=torch.uint8, device="cuda"), 3)
    ii = api.integral_batch(src)
    for _ in range(2):
        api.lbp_detect_batch(cas, ii, 65536, 1.1, 1.0, 4.0, 2)
torch.cuda.synchronize()
print("done", op)
