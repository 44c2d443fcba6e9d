#!/bin/bash
# End-to-end smoke of the drop-in CLIs on a GPU box (tiny epoch counts).
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import time, numpy as np
from mr_gan_b200.mr_gan import dataset, mr_gan, train_gan_folds, _kfold_jobs
from mr_gan_b200.mr_nn import mr_nn
X, y = dataset(modalities=1, synthetic_data=True)            # temperature, D=400 (synthetic MREO shape)
t = time.time(); e = mr_gan(X, y, percentlabeled=16, epochs=3, seed=1, verbose=True); print("mr_gan fp32 3 epochs: err %.4f in %.1fs" % (e, time.time() - t))
t = time.time(); e = mr_gan(X, y, percentlabeled=16, epochs=3, seed=1, precision='tf32'); print("mr_gan tf32 3 epochs: err %.4f in %.1fs" % (e, time.time() - t))
t = time.time(); e = mr_nn(X, y, percentlabeled=16, epochs=5, seed=1, precision='tf32'); print("mr_nn tf32 5 epochs: err %.4f in %.1fs" % (e, time.time() - t))
jobs = _kfold_jobs(X, y, 0, percentlabeled=100)
t = time.time(); errs = train_gan_folds(jobs, epochs=10, seed=3, precision='tf32'); print("6-fold group tf32 10 epochs: errs", np.round(errs, 4), "in %.1fs" % (time.time() - t))
PY
