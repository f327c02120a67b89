#!/usr/bin/env python
"""Divergence / load-balance model of the LBP cascade scan (analysis for DESIGN.md section 6; no GPU needed).

For a c4-like frame (gs_blur r=3 of uniform noise, like bench.py) it takes the oracle's per-window stage
depth at a few scales and replays k_lbp_scan3's schedule -- 64 x 32 window tiles, 16 warps, a warp owning
slots warp, warp+16, ... -- under a simple cost model (one weak classifier for one warp-iteration = 1 unit),
then compares alternatives: dynamic slot hand-out in chunks of 1, 2 or 4 slots.

    python tools/lbp_model.py [width height]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _libs as L  # noqa: E402

CUTS = [1, 2, 3, 4, 6, 9, 13]
FLAT = 8


def warp_cost(depths, stage_n, cuts, nst):
    """cost (weak-classifier warp-iterations) of one warp processing `depths` (stage depth of its windows) with
    lane-per-window groups re-packed at `cuts`, switching to the flat (window, weak) mode at <= FLAT survivors"""
    ends = [c for c in cuts if c < nst] + [nst]
    alive = np.asarray(depths)
    cost = 0.0
    s0 = 0
    for e in ends:
        m = len(alive)
        if m == 0:
            break
        if m <= FLAT and s0 > 0:
            for s in range(s0, nst):
                m = len(alive)
                if m == 0:
                    break
                p = 1
                while p < stage_n[s] and p < 32:
                    p *= 2
                passes = -(-m // (32 // p))
                cost += passes * (1.0 + 0.03 * stage_n[s])        # one vote per lane + the ordered shuffle sum
                alive = alive[alive > s]
            return cost
        iters = -(-m // 32)
        # every lane walks stages s0..e-1 until its window dies; a warp-iteration costs the longest lane
        for i in range(iters):
            chunk = alive[i * 32:(i + 1) * 32]
            last = min(int(chunk.max()), e - 1)                  # deepest stage any lane of this iteration runs
            cost += sum(stage_n[s0:last + 1])
        alive = alive[alive >= e]
        s0 = e
    return cost


def pooled_cost(tile, stage_n, nst, pool_at):
    """hybrid: warps run stages [0, pool_at) on their own slots (lane per window, re-packed at CUTS), then the
    tile's survivors are pooled (one CTA barrier) and finished by all 16 warps together in the flat mode"""
    early = []
    pool = []
    for wp in range(16):
        d = tile[wp::16].ravel()
        early.append(warp_cost(np.minimum(d, pool_at), stage_n, [c for c in CUTS if c < pool_at], pool_at))
        pool.append(d[d >= pool_at])
    alive = np.concatenate(pool)
    tail = 0.0
    for s in range(pool_at, nst):
        m = len(alive)
        if m == 0:
            break
        p = 1
        while p < stage_n[s] and p < 32:
            p *= 2
        tail += -(-m // (16 * (32 // p))) * (1.0 + 0.03 * stage_n[s])
        alive = alive[alive > s]
    return np.array(early), tail


def rebalanced_cost(tile, stage_n, nst, at, barrier_units=1.5):
    """warps run stages [0, at) on their own slots, then ONE CTA barrier: the survivors are dealt out evenly
    (in list order) and every warp finishes its share alone (lane-per-window groups / flat tail, no more barriers)"""
    early, pool = [], []
    for wp in range(16):
        d = tile[wp::16].ravel()
        early.append(warp_cost(np.minimum(d, at), stage_n, [c for c in CUTS if c < at], at))
        pool.append(d[d >= at])
    alive = np.concatenate(pool)
    tails = []
    for wp in range(16):
        share = alive[wp::16]                                   # round-robin deal
        # stages before `at` are already passed: shift so that warp_cost starts at stage `at`
        tails.append(warp_cost(share - at, stage_n[at:], [c - at for c in CUTS if c > at], nst - at) if len(share) else 0.0)
    return np.array(early), np.array(tails), barrier_units


def main():
    w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
    O = L.oracle()
    cas = L.HostCascade()
    a = cas.arrays
    stage_n = [int(v) for v in a["stage_nweaks"]]
    nst = len(stage_n)
    rng = np.random.default_rng(1234)
    noise = rng.integers(0, 256, (h, w)).astype(np.uint8)
    img = np.empty_like(noise); O.gso_blur(L.ptr(img), L.ptr(noise), w, h, 3)
    ii = np.empty((h, w), np.uint32); O.gso_integral(L.ptr(img), w, h, L.ptr(ii))
    print("cascade: %d stages, weaks per stage %s" % (nst, stage_n))
    for scale in (1.0, 1.4641, 2.1436, 3.1384):
        s = float(np.float32(scale))
        win = int(np.float32(24) * np.float32(s))
        nx, ny = (w - win) // 2 + 1, (h - win) // 2 + 1
        depth = np.zeros((ny, nx), np.uint8)
        O.gso_lbp_depth_map(cas.ptr, L.ptr(ii), w, h, s, 2, L.ptr(depth))
        surv = [(depth >= k).mean() for k in range(nst + 1)]
        weaks = sum(stage_n[k] * surv[k] for k in range(nst))
        print("scale %.3f (%dx%d windows): survive stage 1/2/3/4/6: %.3f %.3f %.3f %.3f %.4f; %.2f weak evals per window"
              % (s, nx, ny, surv[1], surv[2], surv[3], surv[4], surv[6], weaks))
        res = {}
        for name, chunk, dynamic in (("static 4 slots/warp (now)", 4, False), ("dynamic, 4-slot chunks", 4, True),
                                     ("dynamic, 2-slot chunks", 2, True), ("dynamic, 1-slot chunks", 1, True)):
            tot_work, tot_crit = 0.0, 0.0
            for ty in range(0, ny - 31, 32):
                for tx in range(0, nx - 63, 64):
                    tile = depth[ty:ty + 32, tx:tx + 64].reshape(64, 32)      # 64 slots of 32 windows (id order)
                    if dynamic:
                        units = [tile[k:k + chunk].ravel() for k in range(0, 64, chunk)]
                        costs = [warp_cost(u, stage_n, CUTS, nst) for u in units]
                        load = np.zeros(16)
                        for c in costs:                                       # greedy hand-out in slot order
                            load[load.argmin()] += c
                    else:
                        load = np.array([warp_cost(tile[wp::16].ravel(), stage_n, CUTS, nst) for wp in range(16)])
                    tot_work += load.sum()
                    tot_crit += load.max() * 16
            res[name] = (tot_work, tot_crit)
        for pool_at in (2, 3, 4, 6):
            tot_work, tot_crit = 0.0, 0.0
            for ty in range(0, ny - 31, 32):
                for tx in range(0, nx - 63, 64):
                    early, tail = pooled_cost(depth[ty:ty + 32, tx:tx + 64].reshape(64, 32), stage_n, nst, pool_at)
                    tot_work += early.sum() + tail * 16
                    tot_crit += (early.max() + tail) * 16
            res["pooled flat tail from stage %d" % pool_at] = (tot_work, tot_crit)
        for at in (1, 2, 3, 4):
            tot_work, tot_crit = 0.0, 0.0
            for ty in range(0, ny - 31, 32):
                for tx in range(0, nx - 63, 64):
                    early, tails, bar = rebalanced_cost(depth[ty:ty + 32, tx:tx + 64].reshape(64, 32), stage_n, nst, at)
                    tot_work += early.sum() + tails.sum()
                    tot_crit += (early.max() + bar + tails.max()) * 16
            res["re-deal survivors once at stage %d" % at] = (tot_work, tot_crit)
        base = res["static 4 slots/warp (now)"][1]
        for name, (work, crit) in res.items():
            print("   %-28s work %.3g  warp-slots held until the CTA retires %.3g  balance %.2f  vs now %.2f"
                  % (name, work, crit, work / crit, crit / base))


if __name__ == "__main__":
    main()
