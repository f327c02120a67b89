"""The BASELINE configs[4] frame pipeline on one rank's shard of frames (SURVEY.md 8d "C5"):

    gs_blur(r=5) -> gs_sobel -> gs_orb_extract(nkps, t)  [on the sobel map]
                             -> gs_integral + gs_lbp_detect(frontalface) [on the sobel map]

Every stage is one batched C-ABI call on device-resident frames; buffers are allocated once per
shard size and reused.  Used by bench.py (--workload c5 and the `shard` section), the sharded GPU
test and gsb_magick-style callers.  No arithmetic happens here.
"""
from . import api
from ._lib import lib, check

C5 = dict(radius=5, nkps=1250, threshold=20, max_rects=4096, scale_factor=1.1, min_scale=1.0, max_scale=4.0, step=2)


class FramePipeline:
    """Preallocated C5 chain for up to `n` frames of h x w on `device`.  fused=True runs blur -> sobel as the one-pass
    gs_b200_blur_sobel_batch (no blurred intermediate: n*h*w bytes less memory and half the HBM traffic of the pair,
    but instruction-bound and ~15 % slower than the two kernels on B200, so it is not the default)."""

    def __init__(self, cascade, n, h, w, device, fused=False, **params):
        import torch
        self.p = dict(C5)
        self.p.update(params)
        self.cascade, self.n, self.h, self.w, self.device = cascade, n, h, w, device
        u8 = dict(dtype=torch.uint8, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        self.fused = fused and hasattr(api, "blur_sobel_batch")
        self.blur = None if self.fused else torch.empty((n, h, w), **u8)
        self.sobel = torch.zeros((n, h, w), **u8)          # border stays 0 like a gs_alloc'ed image
        self.score = torch.zeros((n, h, w), **u8)
        self.ii = torch.empty((n, h, w), **i32)
        self.kps = torch.empty((n, self.p["nkps"], 12), **i32)
        self.kcounts = torch.zeros((n,), **i32)
        self.rects = torch.empty((n, self.p["max_rects"], 4), **i32)
        self.rcounts = torch.zeros((n,), **i32)

    def run(self, frames, lo=0):
        """frames: (m, h, w) uint8 on the device, lo + m <= n; results land in rows [lo, lo+m) of the buffers.
        Asynchronous on torch's current stream."""
        m = frames.shape[0]
        assert lo + m <= self.n
        if m == 0:
            return
        p, L = self.p, lib()
        st = api._stream()
        sl = slice(lo, lo + m)
        sob, sc, ii = self.sobel[sl], self.score[sl], self.ii[sl]
        if self.fused:
            api.blur_sobel_batch(frames, p["radius"], out=sob)
        else:
            api.blur_batch(frames, p["radius"], out=self.blur[sl])
            api.sobel_batch(self.blur[sl], out=sob)
        check(L.gs_b200_orb_extract_batch(api._p(sob), self.w, self.h, m, api._p(sc), api._p(self.kps[sl]),
                                          api._p(self.kcounts[sl]), p["nkps"], p["threshold"], st), "orb_extract_batch")
        api.integral_batch(sob, out=ii)
        check(L.gs_b200_lbp_detect_batch(self.cascade.ptr, api._p(ii), self.w, self.h, m, api._p(self.rects[sl]),
                                         api._p(self.rcounts[sl]), p["max_rects"], p["scale_factor"], p["min_scale"],
                                         p["max_scale"], p["step"], st), "lbp_detect_batch")

    def results(self, m=None):
        m = self.n if m is None else m
        return {"sobel": self.sobel[:m], "kps": self.kps[:m], "kcounts": self.kcounts[:m],
                "rects": self.rects[:m], "rcounts": self.rcounts[:m]}
