#!/bin/bash
# compute-sanitizer memcheck over one small D+G step pair and one tiny epoch in both precisions (run under gpurun).
cd "$(dirname "$0")/.."
cat > /tmp/san.py <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from mr_gan_b200.engine import FoldGroup
from mr_gan_b200.model import init_disc, init_gen
rng = np.random.default_rng(0)
D, B, ntr, nte = 44, 12, 48, 20
for prec in ("fp32", "tf32"):
    with FoldGroup([(D, ntr, nte, 7), (D + 8, ntr, nte, 8)], precision=prec, batch=B) as fg:
        for f, d in enumerate((D, D + 8)):
            fg.set_params(f, 0, init_disc(d, rng)); fg.set_params(f, 1, init_gen(d, rng))
            X = rng.standard_normal((ntr, d)).astype(np.float32); y = (np.arange(ntr) % 6).astype(np.int32)
            fg.load_fold(f, X, y, X[:nte], y[:nte])
        x = rng.standard_normal((B, D)).astype(np.float32)
        print(prec, fg.train_batch_disc(0, x, np.arange(B) % 6, x, rng.standard_normal((B, 100))), fg.train_batch_gen(0, x, rng.standard_normal((B, 100))))
        idx = np.stack([rng.permutation(ntr) for _ in range(2)]).astype(np.int32)
        print(prec, fg.train_epoch(idx, idx, idx)[0], fg.eval(1))
print("SANITIZE RUN DONE")
PY
compute-sanitizer --tool memcheck --error-exitcode 1 python /tmp/san.py 2>&1 | tail -15
