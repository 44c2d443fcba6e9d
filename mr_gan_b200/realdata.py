"""Real-data path of ``dataset()`` (mr_gan.py:23-71): loads the processed MREO pickles written by
processdata.py:91-92 and assembles the per-modality feature rows of mr_gan.py:49-62.

The reference calls librosa 0.5.1 for the contact-microphone features (mr_gan.py:45-47:
``melspectrogram(y, sr=48000, n_mels=128)`` then ``logamplitude(S, ref_power=np.max)``).  librosa is not
available here, so the two functions are restated in numpy from librosa's published definitions
(centered STFT, n_fft=2048, hop=512, periodic Hann window, power spectrogram, Slaney mel scale with
area-normalised triangular filters, 10*log10 relative to the maximum, floor at -80 dB).  The MREO files are
not distributed with the reference, so this path is covered by format / property tests and pinned piecewise against
independent implementations (tests/test_host.py: the STFT against scipy.signal.stft to 1e-10, the mel scale against Slaney's
landmarks, the filters against their definition) -- NOT verified against librosa output itself."""
import os
import pickle
import sys

import numpy as np

from .model import MATERIALS


# ------------------------------------------------------------------ librosa-free log-mel
def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    mel = f / f_sp
    log_t = f >= min_log_hz
    return np.where(log_t, min_log_hz / f_sp + np.log(np.maximum(f, min_log_hz) / min_log_hz) / (np.log(6.4) / 27.0), mel)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp((np.log(6.4) / 27.0) * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr=48000, n_fft=2048, n_mels=128, fmin=0.0, fmax=None):
    """librosa.filters.mel (Slaney scale, norm=1): [n_mels, 1 + n_fft//2]."""
    fmax = sr / 2.0 if fmax is None else fmax
    fft_f = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_f[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    return w * (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]


def melspectrogram(y, sr=48000, n_fft=2048, hop_length=512, n_mels=128):
    """librosa.feature.melspectrogram(y, sr, n_mels) with librosa's defaults: power spectrogram -> mel."""
    y = np.asarray(y, dtype=np.float64)
    y = np.pad(y, n_fft // 2, mode='reflect')
    n_frames = 1 + (len(y) - n_fft) // hop_length
    idx = np.arange(n_fft)[None, :] + hop_length * np.arange(n_frames)[:, None]
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft) / n_fft)          # periodic Hann
    S = np.abs(np.fft.rfft(y[idx] * win, axis=1)) ** 2                        # [frames, bins]
    return mel_filterbank(sr, n_fft, n_mels) @ S.T                            # [n_mels, frames]


def logamplitude(S, amin=1e-10, top_db=80.0):
    """librosa.logamplitude(S, ref_power=np.max): dB relative to the peak, floored at -top_db."""
    S = np.asarray(S, dtype=np.float64)
    log_spec = 10.0 * np.log10(np.maximum(amin, S)) - 10.0 * np.log10(np.maximum(amin, S.max()))
    return np.maximum(log_spec, log_spec.max() - top_db)


# ------------------------------------------------------------------ processed-pickle loader
def processed_path(data_dir, material, forcetempTime, contactmicTime):
    return os.path.join(data_dir, 'processed_0.1sbefore_%s_times_%.2f_%.2f.pkl' % (material, forcetempTime, contactmicTime))


def load_processed(modalities=0, forcetempTime=4, contactmicTime=0.2, leaveObjectOut=False, verbose=False,
                   data_dir='data_processed'):
    """mr_gan.py:23-71 on the files of processdata.py:91 ({object: {force0, force1, temperature, contact: list of lists}})."""
    X, y, objects = [], [], dict()
    for m, material in enumerate(MATERIALS):
        if verbose:
            print('Processing', material)
            sys.stdout.flush()
        with open(processed_path(data_dir, material, forcetempTime, contactmicTime), 'rb') as f:
            allData = pickle.load(f, encoding='latin1')          # Python-2 pickles
        for objName, objData in allData.items():
            if leaveObjectOut:
                objects[objName] = {'x': [], 'y': []}
                X, y = objects[objName]['x'], objects[objName]['y']
            for i in range(len(objData['temperature'])):
                y.append(m)
                temp = list(objData['temperature'][i])
                force = list(objData['force0'][i]) + list(objData['force1'][i])
                mel = []
                if modalities > 2:                                # mr_gan.py:42-47
                    mel = logamplitude(melspectrogram(np.array(objData['contact'][i]), sr=48000, n_mels=128)).flatten().tolist()
                row = {0: force, 1: temp, 2: temp + force, 3: mel, 4: temp + mel, 5: temp + force + mel, 6: force + mel}[modalities]
                X.append(row)
    if leaveObjectOut:
        return {k: {'x': np.array(v['x']), 'y': np.array(v['y'])} for k, v in objects.items()}
    X, y = np.array(X), np.array(y)
    if verbose:
        print('X:', np.shape(X), 'y:', np.shape(y))
    return X, y
