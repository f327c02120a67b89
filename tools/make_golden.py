#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference (oracle/_ref/libgs_ref.so, the
unmodified /root/reference/grayskull.h compiled by oracle/Makefile) in this container.
The GPU box has no /root/reference, so these small fixtures are committed.

  lena_golden.npz     every hot-path op on testdata/lena.pgm (128x128; BASELINE config C1)
  random_golden.npz   a handful of odd-sized random images (ragged widths, tiny sizes)
  next_golden.npz     the SURVEY.md 8(f) rows (gs_match_orb, histogram / Otsu / threshold, gs_filter,
                      gs_match_template / gs_find_best_match) on lena and two synthetic images
  round2_golden.npz   radii beyond 7 (gs_blur / gs_adaptive_threshold r = 8..31, the reference Makefile's
                      `blur 9` / `adaptive 15 5`), gs_compute_orientation for r > 15, and the 8(f) N4 row:
                      gs_blobs (incl. running out of labels), gs_blob_corners, gs_perspective_correct
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _libs as L  # noqa: E402

REF_TREE = os.environ.get("GS_REFERENCE", "/root/reference")


def read_pgm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"P5"
        w, h = map(int, f.readline().split())
        assert int(f.readline()) == 255
        return np.frombuffer(f.read(w * h), np.uint8).reshape(h, w).copy()


def run_all(R, a, cas_ptr, tag, out, orb_nkps=500, lbp=True):
    h, w = a.shape
    d = np.zeros_like(a); R.gs_sobel(L.img(d), L.img(a)); out[tag + "sobel"] = d
    for r in (1, 5, 9):
        d = np.empty_like(a); R.gs_blur(L.img(d), L.img(a), r); out[tag + "blur%d" % r] = d
    d = np.empty_like(a); R.gs_adaptive_threshold(L.img(d), L.img(a), 15, 5); out[tag + "adaptive_15_5"] = d
    d = np.empty_like(a); R.gs_erode(L.img(d), L.img(a)); out[tag + "erode"] = d
    d = np.empty_like(a); R.gs_dilate(L.img(d), L.img(a)); out[tag + "dilate"] = d
    d = np.empty((64, 128), np.uint8); R.gs_resize(L.img(d), L.img(a)); out[tag + "resize_128x64"] = d
    d = np.empty((h // 2, w // 2), np.uint8); R.gs_downsample(L.img(d), L.img(a)); out[tag + "downsample"] = d
    ii = np.empty(a.shape, np.uint32); R.gs_integral(L.img(a), L.ptr(ii)); out[tag + "integral"] = ii
    sm = np.zeros_like(a); k = np.zeros(5000, L.KP_DTYPE)
    n = R.gs_fast(L.img(a), L.img(sm), L.ptr(k), 5000, 20)
    out[tag + "fast_kps"] = k[:n].copy(); out[tag + "fast_scoremap"] = sm
    sm = np.zeros_like(a); k = np.zeros(orb_nkps, L.KP_DTYPE)
    n = R.gs_orb_extract(L.img(a), L.ptr(k), orb_nkps, 20, L.ptr(sm))
    out[tag + "orb_kps"] = k[:n].copy()
    if lbp:
        r = np.zeros(1000, L.RECT_DTYPE)
        n = R.gs_lbp_detect(cas_ptr, L.ptr(ii), w, h, L.ptr(r), 1000, 1.1, 1.0, 4.0, 2)
        out[tag + "lbp_rects"] = r[:n].copy()


def run_next(R, a, tag, out, rng):
    """reference outputs for the 8(f) rows; inputs that are not derivable from `a` are stored too"""
    h, w = a.shape
    hist = np.zeros(256, np.uint32); R.gs_histogram(L.img(a), L.ptr(hist)); out[tag + "hist"] = hist
    t = int(R.gs_otsu_threshold(L.img(a))); out[tag + "otsu"] = np.array(t)
    d = a.copy(); R.gs_threshold(L.img(d), t); out[tag + "threshold_otsu"] = d
    for name in ("sharpen", "emboss", "box", "gaussian", "emboss_norm3", "k5", "k2x4"):
        k, norm = L.filter_kernel(name)
        d = np.zeros_like(a); R.gs_filter(L.img(d), L.img(a), L.img(k), norm); out[tag + "filter_" + name] = d
    th, tw = min(h, 21), min(w, 30)
    y0, x0 = (h - th) // 3, (w - tw) // 2
    tmpl = np.clip(a[y0:y0 + th, x0:x0 + tw].astype(np.int16) + rng.integers(-2, 3, (th, tw)), 0, 255).astype(np.uint8)
    out[tag + "tmpl"] = tmpl
    res = np.zeros((h - th + 1, w - tw + 1), np.uint8)
    R.gs_match_template(L.img(a), L.img(tmpl), L.img(res)); out[tag + "tmatch"] = res
    p = R.gs_find_best_match(L.img(res)); out[tag + "tmatch_best"] = np.array([p.x, p.y])
    # ORB keypoints of the image and of a shifted, slightly noisy copy, matched with the 0.8 ratio test
    b = np.roll(a, (2, 3), axis=(0, 1))
    b = np.clip(b.astype(np.int16) + rng.integers(-2, 3, b.shape), 0, 255).astype(np.uint8)
    out[tag + "shifted"] = b
    ks = []
    for im in (a, b):
        sm = np.zeros_like(im); k = np.zeros(300, L.KP_DTYPE)
        n = R.gs_orb_extract(L.img(im), L.ptr(k), 300, 20, L.ptr(sm)); ks.append(k[:n].copy())
    out[tag + "kps_a"], out[tag + "kps_b"] = ks
    m = np.zeros(300, L.MATCH_DTYPE)
    n = R.gs_match_orb(L.ptr(ks[0]), len(ks[0]), L.ptr(ks[1] if len(ks[1]) else np.zeros(1, L.KP_DTYPE)), len(ks[1]), L.ptr(m), 300, 60.0)
    out[tag + "matches"] = m[:n].copy()
    return n


def run_round2(R, lena, out):
    rng = np.random.default_rng(2028)
    # ---- large radii
    imgs = {"lena": lena, "rag": rng.integers(0, 256, (75, 150)).astype(np.uint8), "w16": L.natural_like(208, 97, 5),
            "tiny": rng.integers(0, 256, (9, 16)).astype(np.uint8)}
    out["radius_tags"] = np.array(list(imgs))
    out["radii"] = np.array([8, 9, 11, 15, 31])
    for tag, a in imgs.items():
        out["radius_img_" + tag] = a
        for r in (8, 9, 11, 15, 31):
            d = np.empty_like(a); R.gs_blur(L.img(d), L.img(a), r); out["blur%d_%s" % (r, tag)] = d
            d = np.empty_like(a); R.gs_adaptive_threshold(L.img(d), L.img(a), r, 5 - r); out["adaptive%d_%s" % (r, tag)] = d
    # ---- orientation beyond r = 15 (float accumulation rounds there)
    big = np.clip(L.natural_like(400, 300, 11).astype(np.int32) + 90, 0, 255).astype(np.uint8)
    out["orient_img"] = big
    xyr, ang = [], []
    for r in (3, 15, 16, 24, 40, 77, 120):
        for _ in range(4):
            x, y = int(rng.integers(r, 400 - r)), int(rng.integers(r, 300 - r))
            xyr.append((x, y, r)); ang.append(R.gs_compute_orientation(L.img(big), x, y, r))
    out["orient_xyr"], out["orient_angle"] = np.array(xyr), np.array(ang, np.float32)
    # ---- blobs / corners / perspective
    bimgs = {"b0": L.binary_like(64, 48, 1), "b1": L.binary_like(200, 120, 2, density=0.35), "b2": L.binary_like(131, 77, 3, density=0.6, smooth=1),
             "b3": (lena > 110).astype(np.uint8) * 255, "b4": np.full((20, 33), 255, np.uint8), "b5": np.zeros((10, 10), np.uint8),
             "b6": L.binary_like(512, 300, 4, density=0.45, smooth=5)}
    snake = np.zeros((41, 64), np.uint8)           # comb / snake shapes: provisional labels merge late and in both directions
    snake[::2, :] = 255; snake[1::4, 0] = 255; snake[3::4, -1] = 255
    bimgs["b7"] = snake
    comb = np.zeros((30, 61), np.uint8); comb[:, ::2] = 255; comb[-1, :] = 255
    bimgs["b8"] = comb
    out["blob_tags"] = np.array(list(bimgs))
    for tag, a in bimgs.items():
        out["blob_img_" + tag] = a
        h, w = a.shape
        for nb in (1000, 7, 1):
            labels = np.full((h, w), 0xABCD, np.uint16)
            blobs = np.zeros(nb, L.BLOB_DTYPE)
            m = R.gs_blobs(L.img(a), L.ptr(labels), L.ptr(blobs), nb)
            out["blob_%s_n%d_labels" % (tag, nb)] = labels
            out["blob_%s_n%d_blobs" % (tag, nb)] = np.array(L.blob_fields(blobs[:m]), np.int64).reshape(m, 8)
            if nb == 1000 and m:
                cs = []
                for i in range(min(m, 12)):
                    c = np.zeros((4, 2), np.uint32)
                    R.gs_blob_corners(L.img(a), L.ptr(labels), L.ptr(blobs[i:i + 1]), L.ptr(c)); cs.append(c)
                out["blob_%s_corners" % tag] = np.array(cs)
    src = L.natural_like(300, 220, 21)
    out["persp_src"] = src
    quads = [[(20, 30), (270, 10), (290, 200), (5, 180)], [(0, 0), (299, 0), (299, 219), (0, 219)], [(250, 200), (10, 190), (30, 15), (280, 40)],
             [(100, 100), (100, 100), (100, 100), (100, 100)], [(0, 0), (1000, 5), (900, 700), (3, 400)]]
    out["persp_quads"] = np.array(quads, np.uint32)
    for qi, q in enumerate(quads):
        for (dw, dh) in ((160, 100), (33, 47), (1, 1), (2, 5)):
            d = np.empty((dh, dw), np.uint8)
            c = np.array(q, np.uint32)
            R.gs_perspective_correct(L.img(d), L.img(src), L.ptr(c)); out["persp_q%d_%dx%d" % (qi, dw, dh)] = d


def main():
    R = L.ref()
    cas = R.ref_frontalface()
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)

    lena = read_pgm(os.path.join(REF_TREE, "testdata", "lena.pgm"))
    out = {"lena": lena, "lena_md5": hashlib.md5(lena.tobytes()).hexdigest()}
    run_all(R, lena, cas, "", out)
    np.savez_compressed(os.path.join(gold, "lena_golden.npz"), **out)
    print("lena: fast %d kps, orb %d kps, lbp %d rects" % (len(out["fast_kps"]), len(out["orb_kps"]), len(out["lbp_rects"])))

    rng = np.random.default_rng(2026)
    out = {}
    shapes = [(33, 29), (64, 48), (100, 37), (48, 160), (256, 64), (130, 131)]
    out["shapes"] = np.array(shapes)
    for i, (w, h) in enumerate(shapes):
        a = L.natural_like(w, h, seed=100 + i) if i % 2 else rng.integers(0, 256, (h, w)).astype(np.uint8)
        out["img%d" % i] = a
        run_all(R, a, cas, "i%d_" % i, out, orb_nkps=200, lbp=(w >= 24 and h >= 24))
    np.savez_compressed(os.path.join(gold, "random_golden.npz"), **out)

    rng = np.random.default_rng(2027)
    out = {"tags": np.array(["lena_", "nat_", "rnd_"])}
    imgs = {"lena_": lena, "nat_": L.natural_like(200, 120, 7), "rnd_": rng.integers(0, 256, (50, 77)).astype(np.uint8)}
    for tag, a in imgs.items():
        out[tag + "img"] = a
        n = run_next(R, a, tag, out, rng)
        print("%s %d matches, otsu %d, best %s" % (tag, n, int(out[tag + "otsu"]), out[tag + "tmatch_best"]))
    np.savez_compressed(os.path.join(gold, "next_golden.npz"), **out)
    out = {}
    run_round2(R, lena, out)
    np.savez_compressed(os.path.join(gold, "round2_golden.npz"), **out)
    print("round2: %d arrays" % len(out))
    print("wrote", gold)


if __name__ == "__main__":
    main()
