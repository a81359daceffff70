"""Summarises a B200_PCG_TRACE dump (phase time stamps of the persistent PCG kernel, ns).
   python tools/pcg_trace.py trace.bin [num_ctas]"""
import sys
import numpy as np
raw = np.fromfile(sys.argv[1], dtype=np.uint64)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 148
t = raw.reshape(n, -1, 8).astype(np.float64)
iters = t.shape[1]
valid = [i for i in range(iters) if (t[:, i, 7] > 0).all() and (t[:, i, 0] > 0).all()]
print("iterations with complete stamps:", len(valid))
names = ["tiles(warp0)", "big+prime+flush", "barrier1", "alpha,x,r", "precond+sums", "barrier2", "tests,beta,p"]
rows = []
for i in valid[2:]:
    d = np.diff(t[:, i, :], axis=1) / 1e3  # us, [cta][7]
    rows.append(d)
d = np.stack(rows)  # [iter][cta][7]
print("%-18s %8s %8s %8s" % ("phase", "median", "mean", "max"))
for k, nm in enumerate(names):
    print("%-18s %8.2f %8.2f %8.2f" % (nm, np.median(d[:, :, k]), d[:, :, k].mean(), d[:, :, k].max()))
it_time = np.diff(t[0, valid, 0]) / 1e3
print("iteration time (cta 0): median %.2f us, mean %.2f us" % (np.median(it_time), it_time.mean()))
prod_end = t[:, valid[2:], 2] - t[:, valid[2:], 0]
print("product+flush per CTA: min %.2f median %.2f max %.2f us (spread across CTAs is waited out at barrier 1)" %
      (prod_end.min(axis=0).mean() / 1e3, np.median(prod_end, axis=0).mean() / 1e3, prod_end.max(axis=0).mean() / 1e3))
start_skew = t[:, valid[2:], 0] - t[:, valid[2:], 0].min(axis=0)
print("start skew across CTAs: mean max %.2f us" % (start_skew.max(axis=0).mean() / 1e3))
