#!/usr/bin/env python
"""Decode the persistent conv kernel's debug timeline (ICP_TIMELINE=<file>): per-K-block clock64 stamps of CTA 0.
rows: 0 A-producer passed empty wait, 1 W-producer passed empty wait, 2 MMA saw full, 3 MMA issued+committed,
4 epilogue saw tmem_full, 5 epilogue done, 6 epilogue started waiting."""
import sys
import numpy as np
d = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, 512)
num_kb = int(sys.argv[2]) if len(sys.argv) > 2 else 16
t0 = d[d > 0].min()
r = np.where(d > 0, d - t0, -1)
n = int((r[2] >= 0).sum())
print("k-blocks recorded:", n, " tiles:", int((r[4] >= 0).sum()))
print("%4s %9s %9s %9s %9s | %7s %7s" % ("kb", "A.issue", "W.issue", "MMA.full", "MMA.done", "dMMA", "full-A"))
for i in range(min(n, 80)):
    print("%4d %9d %9d %9d %9d | %7d %7d" % (i, r[0, i], r[1, i], r[2, i], r[3, i], r[2, i] - (r[2, i - 1] if i else 0),
                                          r[2, i] - r[0, i]))
dm = np.diff(r[2, :n])
print("MMA full->full per kb: median %d  mean %.0f  p90 %d" % (np.median(dm), dm.mean(), np.percentile(dm, 90)))
print("MMA issue span per kb (full->done): median %d" % np.median(r[3, :n] - r[2, :n]))
print("load latency (A issue -> MMA sees full): median %d" % np.median(r[2, :n] - r[0, :n]))
nt = int((r[4] >= 0).sum())
for t in range(min(nt, 12)):
    print("tile %2d: epi wait-start %8d  got full %8d  done %8d  (epilogue %d cycles)" % (t, r[6, t], r[4, t], r[5, t], r[5, t] - r[4, t]))

if r.shape[0] >= 10 and (r[7] >= 0).any():
    ns = int((r[7] >= 0).sum())
    print("epilogue units (rows 7/8/9) — halo: TMEM read done, staging written+barrier, copy-out issued;")
    print("                                persistent: arithmetic done, staging free (wait_read+barrier), TMA store issued:")
    for q in range(min(ns, 24)):
        print("  sub %3d: tmem %8d  staged %8d (+%d)  stored %8d (+%d)" % (q, r[7, q], r[8, q], r[8, q] - r[7, q], r[9, q], r[9, q] - r[8, q]))

if r.shape[0] >= 14 and (r[13] >= 0).any():
    print("persistent epilogue chunk phases: ldwait-done | +convert (7) | +free (8) | +STS (10) | +fence (11) || store warp: issued (9) - fence | read done (12) - issued")
    ns = int((r[13] >= 0).sum())
    for q in range(min(ns, 24)):
        t = r[13, q]
        print("  chunk %3d: t=%8d  %5d %5d %5d %5d || %5d %5d" % (q, t, r[7, q] - t, r[8, q] - r[7, q], r[10, q] - r[8, q],
                                                      r[11, q] - r[10, q], r[9, q] - r[11, q], r[12, q] - r[9, q]))
