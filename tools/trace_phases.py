#!/usr/bin/env python3
"""Reads the per-CTA phase timeline of the profiling build (M1_TRACE, tools/experiments) and reports how the
colour and block phases of the CTAs resident on one SM overlap in time.
usage: trace_phases.py trace.bin"""
import sys
import numpy as np
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 4)
a = a[a[:, 3] != 0]
sm = (a[:, 0] >> np.uint64(48)).astype(np.int64)
t0 = (a[:, 0] & np.uint64((1 << 48) - 1)).astype(np.int64)
t1, t2, t3 = (a[:, i].astype(np.int64) for i in (1, 2, 3))
base = t0.min()
t0 -= base; t1 -= base; t2 -= base; t3 -= base
print(f"CTAs {len(a)}, SMs {len(np.unique(sm))}, kernel span {t3.max() / 1e3:.1f} us")
print(f"mean ns: colour {np.mean(t1 - t0):.0f}  blocks {np.mean(t2 - t1):.0f}  tail {np.mean(t3 - t2):.0f}  life {np.mean(t3 - t0):.0f}")
for lo, hi in ((0.0, 0.1), (0.1, 0.3), (0.3, 0.6), (0.6, 0.95)):
    T0, T1 = t3.max() * lo, t3.max() * hi
    hist = np.zeros(16)
    res = np.zeros(16)
    for s in np.unique(sm):
        m = sm == s
        ev = []
        for b, e in zip(t0[m], t1[m]): ev += [(b, 1, 0), (e, -1, 0)]
        for b, e in zip(t0[m], t3[m]): ev += [(b, 0, 1), (e, 0, -1)]
        ev.sort()
        c = r = 0; last = 0
        for t, dc, dr in ev:
            a0, a1 = max(last, T0), min(t, T1)
            if a1 > a0: hist[c] += a1 - a0; res[r] += a1 - a0
            c += dc; r += dr; last = t
    hist /= hist.sum(); res /= res.sum()
    print(f"window {lo:.2f}-{hi:.2f} of the launch: time share by number of resident CTAs in the colour phase: " +
          " ".join(f"{i}:{100 * h:.0f}%" for i, h in enumerate(hist[:9])) + "   | resident CTAs: " +
          " ".join(f"{i}:{100 * h:.0f}%" for i, h in enumerate(res[:9]) if h > 0.005))
