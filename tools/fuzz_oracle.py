#!/usr/bin/env python
"""Time-boxed differential fuzzing of the oracle restatement (oracle/gs_oracle.c) against the REAL reference
(oracle/_ref/libgs_ref.so): wider parameter ranges than tests/test_oracle.py -- radii larger than the image,
1-pixel-wide images, extreme thresholds / norms / scale ladders.  TEST INFRASTRUCTURE; needs /root/reference
at build time, so it runs in the authoring container only.

    python tools/fuzz_oracle.py [seconds] [seed]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _libs as L  # noqa: E402


def image(rng, w, h):
    kind = rng.integers(0, 5)
    if kind == 0:
        return rng.integers(0, 256, (h, w)).astype(np.uint8)
    if kind == 1:
        return L.natural_like(w, h, int(rng.integers(0, 1 << 30)))
    if kind == 2:
        return np.full((h, w), int(rng.integers(0, 256)), np.uint8)
    if kind == 3:
        return (rng.integers(0, 2, (h, w)) * 255).astype(np.uint8)
    a = rng.integers(0, 256, (h, w)).astype(np.uint8)
    a[rng.random((h, w)) < 0.5] = int(rng.integers(0, 256))
    return a


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
    rng = np.random.default_rng(seed)
    R, O = L.ref(), L.oracle()
    cas, ref_c = L.HostCascade(), R.ref_frontalface()
    counts = {}
    t0 = time.time()

    def hit(name):
        counts[name] = counts.get(name, 0) + 1

    def same(name, x, y, ctx):
        hit(name)
        ok = (x.tobytes() == y.tobytes()) if hasattr(x, "tobytes") else (x == y)
        if not ok:
            print("MISMATCH", name, ctx)
            sys.exit(1)

    while time.time() - t0 < budget:
        w, h = int(rng.integers(1, 97)), int(rng.integers(1, 97))
        if rng.random() < 0.15:
            w, h = (1, h) if rng.random() < 0.5 else (w, 1)
        a = image(rng, w, h)
        # box filters: radii up to far beyond the image
        r = int(rng.choice([0, 1, 2, 3, 5, 7, 8, 15, 40, 200]))
        d = np.empty_like(a); e = np.empty_like(a)
        R.gs_blur(L.img(d), L.img(a), r); O.gso_blur(L.ptr(e), L.ptr(a), w, h, r); same("blur", d, e, (w, h, r))
        c = int(rng.integers(-300, 301))
        R.gs_adaptive_threshold(L.img(d), L.img(a), r, c); O.gso_adaptive_threshold(L.ptr(e), L.ptr(a), w, h, r, c)
        same("adaptive", d, e, (w, h, r, c))
        for dil in (0, 1):
            (R.gs_dilate if dil else R.gs_erode)(L.img(d), L.img(a)); O.gso_morph(L.ptr(e), L.ptr(a), w, h, dil)
            same("morph", d, e, (w, h, dil))
        if w >= 3 and h >= 3:
            d[:] = 9; e[:] = 9
            R.gs_sobel(L.img(d), L.img(a)); O.gso_sobel(L.ptr(e), L.ptr(a), w, h); same("sobel", d, e, (w, h))
        if w >= 2 and h >= 2:
            d2 = np.empty((h // 2, w // 2), np.uint8); e2 = np.empty_like(d2)
            R.gs_downsample(L.img(d2), L.img(a)); O.gso_downsample(L.ptr(e2), L.ptr(a), w, h); same("downsample", d2, e2, (w, h))
        dw, dh = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        d3 = np.empty((dh, dw), np.uint8); e3 = np.empty_like(d3)
        R.gs_resize(L.img(d3), L.img(a)); O.gso_resize(L.ptr(e3), dw, dh, L.ptr(a), w, h); same("resize", d3, e3, (w, h, dw, dh))
        i1 = np.empty(a.shape, np.uint32); i2 = np.empty_like(i1)
        R.gs_integral(L.img(a), L.ptr(i1)); O.gso_integral(L.ptr(a), w, h, L.ptr(i2)); same("integral", i1, i2, (w, h))
        # histogram / otsu / threshold
        h1 = np.zeros(256, np.uint32); h2 = np.zeros(256, np.uint32)
        R.gs_histogram(L.img(a), L.ptr(h1)); O.gso_histogram(L.ptr(a), w, h, L.ptr(h2)); same("histogram", h1, h2, (w, h))
        same("otsu", int(R.gs_otsu_threshold(L.img(a))), int(O.gso_otsu_threshold(L.ptr(a), w, h)), (w, h))
        t = int(rng.integers(0, 256))
        x, y = a.copy(), a.copy()
        R.gs_threshold(L.img(x), t); O.gso_threshold(L.ptr(y), w, h, t); same("threshold", x, y, (w, h, t))
        # generic filter: random kernels, sizes 1..7, norms from 1 to beyond 2^31
        kw, kh = int(rng.integers(1, 8)), int(rng.integers(1, 8))
        k = rng.integers(-128, 128, (kh, kw)).astype(np.int8).view(np.uint8)
        if rng.random() < 0.5:
            k = (rng.integers(-3, 4, (kh, kw))).astype(np.int8).view(np.uint8)
        norm = int(rng.choice([1, 2, 3, 9, 16, 255, 256, 1000, 65536, 20000000, 3000000000, int(rng.integers(1, 1 << 32))]))
        R.gs_filter(L.img(d), L.img(a), L.img(np.ascontiguousarray(k)), norm)
        O.gso_filter(L.ptr(e), L.ptr(a), w, h, L.ptr(np.ascontiguousarray(k)), kw, kh, norm)
        same("filter", d, e, (w, h, kw, kh, norm))
        # template matching
        tw, th = int(rng.integers(1, w + 1)), int(rng.integers(1, h + 1))
        tm = image(rng, tw, th) if rng.random() < 0.5 else np.ascontiguousarray(a[:th, :tw])
        r1 = np.zeros((h - th + 1, w - tw + 1), np.uint8); r2 = np.zeros_like(r1)
        R.gs_match_template(L.img(a), L.img(tm), L.img(r1)); O.gso_match_template(L.ptr(a), w, h, L.ptr(tm), tw, th, L.ptr(r2))
        same("match_template", r1, r2, (w, h, tw, th))
        p = R.gs_find_best_match(L.img(r1))
        same("best_match", p.y * r1.shape[1] + p.x, O.gso_find_best_match(L.ptr(r2), r2.shape[1], r2.shape[0]), (w, h, tw, th))
        # FAST / ORB (the reference needs h >= 7 and w >= 7 for its unsigned loop bounds)
        if w >= 7 and h >= 7:
            t = int(rng.choice([0, 1, 5, 20, 60, 128, 254, 255, 300, 1000]))
            cap = int(rng.choice([1, 2, 17, 300, 5000]))
            sm0 = (rng.integers(0, 256, a.shape) * (rng.random(a.shape) < 0.1)).astype(np.uint8) if rng.random() < 0.5 else np.zeros_like(a)
            s1, s2 = sm0.copy(), sm0.copy()
            k1 = np.zeros(cap, L.KP_DTYPE); k2 = np.zeros(cap, L.KP_DTYPE)
            n1 = R.gs_fast(L.img(a), L.img(s1), L.ptr(k1), cap, t)
            n2 = O.gso_fast(L.ptr(a), w, h, L.ptr(s2), w, h, L.ptr(k2), cap, t)
            same("fast", k1[:n1], k2[:n2], (w, h, t, cap)); same("fast_map", s1, s2, (w, h, t, cap))
        if w >= 36 and h >= 36:
            nk = int(rng.choice([1, 10, 200, 1250]))
            t = int(rng.choice([5, 20, 40]))
            s1, s2 = np.zeros_like(a), np.zeros_like(a)
            k1 = np.zeros(nk, L.KP_DTYPE); k2 = np.zeros(nk, L.KP_DTYPE)
            n1 = R.gs_orb_extract(L.img(a), L.ptr(k1), nk, t, L.ptr(s1))
            n2 = O.gso_orb_extract(L.ptr(a), w, h, L.ptr(k2), nk, t, L.ptr(s2))
            same("orb", k1[:n1], k2[:n2], (w, h, nk, t))
            if n1:
                b = image(rng, w, h)
                k3 = np.zeros(nk, L.KP_DTYPE); s3 = np.zeros_like(b)
                n3 = R.gs_orb_extract(L.img(b), L.ptr(k3), nk, t, L.ptr(s3))
                md = float(rng.choice([0.0, 10.0, 60.0, 64.5, 255.0, 1000.0]))
                mm = int(rng.choice([1, 5, 300]))
                m1 = np.zeros(mm, L.MATCH_DTYPE); m2 = np.zeros(mm, L.MATCH_DTYPE)
                kk = k3 if n3 else np.zeros(1, L.KP_DTYPE)
                c1 = R.gs_match_orb(L.ptr(k1), n1, L.ptr(kk), n3, L.ptr(m1), mm, md)
                c2 = O.gso_match_orb(L.ptr(k1), n1, L.ptr(kk), n3, L.ptr(m2), mm, md)
                same("match_orb", m1[:c1], m2[:c2], (w, h, nk, md, mm))
        # LBP cascade: windows and detection over odd ladders
        if w >= 24 and h >= 24:
            for _ in range(8):
                s = float(np.float32(rng.uniform(1.0, min(w, h) / 24.0)))
                ww = int(np.float32(24.0) * np.float32(s))
                if ww > w or ww > h:
                    continue
                x, y = int(rng.integers(0, w - ww + 1)), int(rng.integers(0, h - ww + 1))
                same("lbp_window", int(R.gs_lbp_window(ref_c, L.ptr(i1), w, h, x, y, s)),
                     int(O.gso_lbp_window(cas.ptr, L.ptr(i2), w, h, x, y, s)), (w, h, x, y, s))
            sf = float(np.float32(rng.choice([1.05, 1.1, 1.2, 1.5, 2.0])))
            mn = float(np.float32(rng.choice([1.0, 1.3, 2.0])))
            mx = float(np.float32(rng.choice([1.0, 2.0, 4.0, 10.0])))
            st = int(rng.integers(1, 6))
            mr = int(rng.choice([1, 3, 1000]))
            q1 = np.zeros(mr, L.RECT_DTYPE); q2 = np.zeros(mr, L.RECT_DTYPE)
            n1 = R.gs_lbp_detect(ref_c, L.ptr(i1), w, h, L.ptr(q1), mr, sf, mn, mx, st)
            n2 = O.gso_lbp_detect(cas.ptr, L.ptr(i2), w, h, L.ptr(q2), mr, sf, mn, mx, st)
            same("lbp_detect", q1[:n1], q2[:n2], (w, h, sf, mn, mx, st, mr))
        # round-2 rows: blobs with and without label overflow, blob corners, perspective warps
        # (degenerate quads included), orientation with radii whose float sums round
        b = L.binary_like(w, h, int(rng.integers(0, 1 << 30)), density=float(rng.uniform(0.1, 0.9)), smooth=int(rng.integers(0, 4))) \
            if rng.random() < 0.8 else a
        for nb in (2000, int(rng.integers(1, 40))):
            l1 = np.full(b.shape, 0x5555, np.uint16); l2 = l1.copy()
            b1 = np.zeros(nb, L.BLOB_DTYPE); b2 = np.zeros(nb, L.BLOB_DTYPE)
            m1 = R.gs_blobs(L.img(b), L.ptr(l1), L.ptr(b1), nb)
            m2 = O.gso_blobs(L.ptr(b), w, h, L.ptr(l2), L.ptr(b2), nb)
            same("blobs_labels", l1, l2, (w, h, nb)); same("blobs", L.blob_fields(b1[:m1]), L.blob_fields(b2[:m2]), (w, h, nb))
            for j in range(min(m1, 3)):
                c1, c2 = np.zeros((4, 2), np.uint32), np.zeros((4, 2), np.uint32)
                R.gs_blob_corners(L.img(b), L.ptr(l1), L.ptr(b1[j:j + 1]), L.ptr(c1))
                O.gso_blob_corners(L.ptr(b), w, h, L.ptr(l2), L.ptr(b2[j:j + 1]), L.ptr(c2))
                same("blob_corners", c1, c2, (w, h, nb, j))
        pw, ph = int(rng.integers(1, 80)), int(rng.integers(1, 80))
        quad = rng.integers(0, max(w, h) + 20, (4, 2)).astype(np.uint32)
        if rng.random() < 0.2:
            quad[int(rng.integers(0, 4))] = quad[int(rng.integers(0, 4))]          # coincident corners
        p1, p2 = np.empty((ph, pw), np.uint8), np.empty((ph, pw), np.uint8)
        R.gs_perspective_correct(L.img(p1), L.img(a), L.ptr(quad)); O.gso_perspective_correct(L.ptr(p2), pw, ph, L.ptr(a), w, h, L.ptr(quad))
        same("perspective", p1, p2, (w, h, pw, ph, quad.tolist()))
        if w >= 3 and h >= 3:                                     # the reference asserts r <= x < w - r, r <= y < h - r
            ro = int(rng.integers(1, (min(w, h) - 1) // 2 + 1))
            ox, oy = int(rng.integers(ro, w - ro)), int(rng.integers(ro, h - ro))
            same("orientation", np.float32(R.gs_compute_orientation(L.img(a), ox, oy, ro)).tobytes(),
                 np.float32(O.gso_compute_orientation(L.ptr(a), w, h, ox, oy, ro)).tobytes(), (w, h, ox, oy, ro))
    print("no mismatch in %.0f s, seed %d:" % (time.time() - t0, seed), " ".join("%s=%d" % kv for kv in sorted(counts.items())))


if __name__ == "__main__":
    main()
