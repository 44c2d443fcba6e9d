import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mr_gan_b200.engine import FoldGroup
from oracle import make_golden, philox
D, B = 100, 50
k = philox.fold_key(3, 1); key = (k[1] << 32) | k[0]
pD, pG, steps = make_golden.case_inputs(D, B, 31, 1)
s = steps[0]
res = {}
for prec in ("fp32", "tf32"):
    with FoldGroup([(D, 100, 40, key)], precision=prec, batch=B) as fg:
        fg.set_params(0, 0, pD); fg.set_params(0, 1, pG)
        fg.train_batch_disc(0, s['x_lab'], s['labels'], s['x_unl'], s['z_d'])
        res[prec] = dict(dlg=fg.debug_buffer(0, 31, 150, 6), h5=fg.debug_buffer(0, 15, 150, 250), dz5=fg.debug_buffer(0, 25, 150, 250),
                         h4=fg.debug_buffer(0, 14, 150, 250), dz4=fg.debug_buffer(0, 24, 150, 250))
for prec in res:
    r = res[prec]
    ref5 = (r['dlg'].astype(np.float64) @ pD[10].astype(np.float64).T) * (r['h5'] > 0)
    e = np.abs(r['dz5'] - ref5)
    print(prec, "dz5 vs numpy: max err %.3e scale %.3e" % (e.max(), np.abs(ref5).max()))
    bad = np.argwhere(e > 1e-2 * np.abs(ref5).max())
    print("  bad count", len(bad), "rows", np.unique(bad[:, 0])[:20], "cols", np.unique(bad[:, 1])[:20])
    for (i, j) in bad[:8]:
        print("   [%d,%d] got %.5e want %.5e h5 %.4e" % (i, j, r['dz5'][i, j], ref5[i, j], r['h5'][i, j]))
    ref4 = (r['dz5'].astype(np.float64) @ pD[8].astype(np.float64).T) * (r['h4'] > 0)
    e = np.abs(r['dz4'] - ref4)
    print(prec, "dz4 vs numpy(from its own dz5): max err %.3e scale %.3e" % (e.max(), np.abs(ref4).max()))
