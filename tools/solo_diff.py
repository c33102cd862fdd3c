import numpy as np
a = np.load("/root/repo/gpurun_out/solo_base.npz"); b = np.load("/root/repo/gpurun_out/solo_nosolo.npz")
for k in a.files:
    d = (a[k] != b[k]).any(-1)
    print(k, "differing pixels", int(d.sum()), "frames", int(d.any((1, 2)).sum()))
    if d.any():
        f = int(np.argmax(d.any((1, 2)))); ys, xs = np.nonzero(d[f])
        print("  frame", f, "y", ys.min(), ys.max(), "x", xs.min(), xs.max(), "base", a[k][f, ys[0], xs[0]], "nosolo", b[k][f, ys[0], xs[0]])
        bins = sorted({(int(y) // 8, int(x) // 32) for y, x in zip(ys, xs)}); print("  coarse bins (row, col):", bins[:20])
