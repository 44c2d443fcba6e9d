"""Synthetic data of the MREO feature shape.

The MREO dataset is not in the reference repo (data_processed/ is empty) and there is no
network, so throughput and parity runs use this generator.  It mirrors ``dataset()``
(mr_gan.py:23-71): 6 materials x 12 objects x 100 pokes = 7200 rows; feature blocks and
their order follow mr_gan.py:49-62; widths follow processdata.py:11-13 (100 Hz force /
temperature windows, 48 kHz contact-mic windows -> 128 mels x (1 + floor(48000*C/512)) frames).

Value model: x = 0.15 * c_class(t) + 0.2 * o_object(t) + eps, with smooth unit-variance random curves per
block and eps ~ N(0,1) iid.  The amplitudes are chosen so that accuracy is neither chance nor 100 %: a
linear classifier reaches ~93 % on a 6-fold split with all labels, ~86 % with 10 % of the labels and ~85 %
leave-one-object-out at D=1200 (the paper reports 95 % / 88 % for force+temperature).  (SURVEY.md 8(d)
suggested 1 : 0.6 : 0.8, which is separable at 100 % by any classifier at these widths.)"""
import numpy as np

from .model import MATERIALS

OBJECTS_PER_MATERIAL = 12
POKES_PER_OBJECT = 100
N_MELS = 128
CLASS_AMPLITUDE = 0.15
OBJECT_AMPLITUDE = 0.2


def mel_frames(contactmicTime):
    return 1 + int(48000 * contactmicTime) // 512      # librosa melspectrogram hop 512, centred


def block_widths(modalities, forcetempTime=4, contactmicTime=0.2):
    """Feature blocks of one row, in the concatenation order of mr_gan.py:49-62."""
    ft = int(round(100 * forcetempTime))
    mel = N_MELS * mel_frames(contactmicTime)
    force, temp, mic = [("force0", ft), ("force1", ft)], [("temperature", ft)], [("contact", mel)]
    return {0: force, 1: temp, 2: temp + force, 3: mic, 4: temp + mic, 5: temp + force + mic, 6: force + mic}[modalities]


def feature_width(modalities, forcetempTime=4, contactmicTime=0.2):
    return sum(w for _, w in block_widths(modalities, forcetempTime, contactmicTime))


def _smooth_curves(rng, n, length):
    """n smooth zero-mean unit-variance random curves of the given length."""
    w = max(1, length // 20)
    walk = np.cumsum(rng.standard_normal((n, length + w - 1)), axis=1)
    if w > 1:
        c = np.cumsum(np.concatenate([np.zeros((n, 1)), walk], axis=1), axis=1)
        walk = (c[:, w:] - c[:, :-w]) / w
    walk = walk[:, :length]
    walk -= walk.mean(axis=1, keepdims=True)
    sd = walk.std(axis=1, keepdims=True)
    return walk / np.where(sd == 0, 1.0, sd)


def synthetic_dataset(modalities=0, forcetempTime=4, contactmicTime=0.2, leaveObjectOut=False, seed=0,
                      pokes=POKES_PER_OBJECT, dtype=np.float64):
    """Same return convention as ``dataset()`` (mr_gan.py:64-71)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    blocks = block_widths(modalities, forcetempTime, contactmicTime)
    K, J = len(MATERIALS), OBJECTS_PER_MATERIAL
    n = K * J * pokes
    X = np.empty((n, sum(w for _, w in blocks)), dtype=dtype)
    y = np.repeat(np.arange(K), J * pokes)
    obj = np.repeat(np.arange(K * J), pokes)
    o = 0
    for _, w in blocks:
        cls = _smooth_curves(rng, K, w)
        ob = _smooth_curves(rng, K * J, w)
        X[:, o:o + w] = CLASS_AMPLITUDE * cls[y] + OBJECT_AMPLITUDE * ob[obj] + rng.standard_normal((n, w))
        o += w
    if not leaveObjectOut:
        return X, y
    objects = {}
    for j in range(K * J):
        rows = obj == j
        objects["%s_%02d" % (MATERIALS[j // J], j % J)] = {"x": X[rows], "y": y[rows]}
    return objects
