"""Prints the per-step loss error of the tf32 path vs the committed golden vectors / fp32 path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mr_gan_b200.engine import FoldGroup
from oracle import make_golden
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for name in ("gan_steps_D36", "gan_steps_D30_B8", "gan_steps_D1200"):
    g = np.load(os.path.join(G, name + ".npz"))
    D, B, n_pairs = int(g['D']), int(g['B']), int(g['n_pairs'])
    key = (int(g['key'][1]) << 32) | int(g['key'][0])
    pD, pG, steps = make_golden.case_inputs(D, B, int(g['seed']), n_pairs)
    for prec in ("fp32", "tf32"):
        with FoldGroup([(D, max(B, 60), 12, key)], precision=prec, batch=B) as fg:
            fg.set_params(0, 0, pD); fg.set_params(0, 1, pG)
            for i, s in enumerate(steps):
                ll, lu, te = fg.train_batch_disc(0, s['x_lab'], s['labels'], s['x_unl'], s['z_d'])
                lg = fg.train_batch_gen(0, s['x_unl2'], s['z_g'])
                w = g['losses'][i]
                print("%s %s step %d rel err: lab %.2e unl %.2e gen %.2e  (err %s)" % (name, prec, i, abs(ll - w[0]) / w[0], abs(lu - w[1]) / w[1], abs(lg - w[3]) / w[3], te - w[2]))
