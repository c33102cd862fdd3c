import numpy as np
a = np.load("/root/repo/gpurun_out/solo2_base.npz"); b = np.load("/root/repo/gpurun_out/solo2_nosolo.npz")
for k in a.files:
    d = (a[k] != b[k]).any(-1)
    print(k, a[k].shape, "differing pixels", int(d.sum()))
    if d.any():
        ys, xs = np.nonzero(d); print("  y", ys.min(), ys.max(), "x", xs.min(), xs.max(), "base", a[k][ys[0], xs[0]], "nosolo", b[k][ys[0], xs[0]])
        print("  coarse bins (row, col):", sorted({(int(y) // 8, int(x) // 32) for y, x in zip(ys, xs)})[:24])
