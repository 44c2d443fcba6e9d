"""ORACLE (test infrastructure only) -- float64 numpy restatement of mr-gan's
semi-supervised feature-matching GAN training step and of mr_nn's supervised step.

PARITY UNPINNED: the reference (Healthcare-Robotics/mr-gan) ships no tests,
golden vectors or fixtures, seeds nothing (mr_gan.py:74-75), and its arithmetic
lives in un-vendored Keras 2.0.9 / Theano 0.9.0 (README.md:41-47) which cannot
be installed here (Python 2.7 only, no network).  This oracle restates the
algorithm from the reference's call sites (cited per function) plus the
published Keras-2.0.9 semantics of the layers/optimizer it calls (listed under
"External semantics" below).  It is self-checked by central finite differences
and cross-checked by an independent torch-autograd twin (oracle/torch_twin.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product
(``mr_gan_b200``) never does.

External semantics (Keras 2.0.9, restated from upstream knowledge)
------------------------------------------------------------------
* ``Dense``: ``y = act(x @ W + b)``, W[in,out] Glorot-uniform
  ``U(+-sqrt(6/(in+out)))``, b = 0.
* ``GaussianNoise(s)``: ``x + N(0, s^2)`` in the training phase, identity in the
  test phase; a fresh draw per layer *application*.
* ``BatchNormalization(epsilon=2e-5)`` in training phase: batch mean, biased
  batch variance, ``gamma*(x-mu)/sqrt(var+eps)+beta``; gamma=1, beta=0 initially.
  Its moving averages are never read on this path.
* ``K.logsumexp`` = stabilised log-sum-exp; ``K.softplus`` = log(1+e^x).
* ``Adam.get_updates``: ``t = iterations+1``; ``lr_t = lr*sqrt(1-b2^t)/(1-b1^t)``;
  ``m = b1 m + (1-b1) g``; ``v = b2 v + (1-b2) g^2``; ``p -= lr_t*m/(sqrt(v)+eps)``.
  ONE Adam object serves both get_updates calls (mr_gan.py:165-167) so the D
  step and the G step share ``iterations`` (D step k sees t=2k-1, G step k
  sees t=2k).  ``shared_t=False`` gives each net its own counter instead.
"""
import numpy as np

# ---- constants: mr_gan.py:77-84, 110-128, 165 ----
NOISE_SIZE = 100                       # mr_gan.py:77
BATCH_GAN = 50                         # mr_gan.py:78
UNLABELED_WEIGHT = 1.0                 # mr_gan.py:79
K_CLASSES = 6                          # mr_gan.py:80
G_HIDDEN = 500                         # mr_gan.py:111,113
BN_EPS = 2e-5                          # mr_gan.py:112
D_WIDTHS = (1000, 500, 250, 250, 250)  # mr_gan.py:119-127
D_SIGMAS = (0.3, 0.5, 0.5, 0.5, 0.5)   # mr_gan.py:118-126
GAN_LR, GAN_B1, GAN_B2, GAN_EPS = 6e-4, 0.5, 0.999, 1e-8   # mr_gan.py:165 + Keras defaults
NN_LR, NN_B1, NN_B2, NN_EPS = 1e-3, 0.9, 0.999, 1e-8       # mr_nn.py:114 ('adam' defaults)
BATCH_NN = 20                          # mr_nn.py:117


def softplus(x):
    return np.logaddexp(0.0, x)


def sigmoid(x):
    return 0.5 * (1.0 + np.tanh(0.5 * x))


def logsumexp(x):
    m = x.max(axis=1, keepdims=True)
    return (m + np.log(np.exp(x - m).sum(axis=1, keepdims=True)))[:, 0]


def softmax(x):
    e = np.exp(x - x.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


# ------------------------------------------------------------------ parameters
def disc_shapes(D, K=K_CLASSES):
    """[W1,b1,...,W6,b6] shapes of the discriminator (mr_gan.py:117-128)."""
    dims = (D,) + D_WIDTHS + (K,)
    out = []
    for i in range(6):
        out += [(dims[i], dims[i + 1]), (dims[i + 1],)]
    return out


def gen_shapes(D):
    """[W1,b1,gamma,beta,W2,b2,W3,b3] shapes of the generator (mr_gan.py:110-114)."""
    return [(NOISE_SIZE, G_HIDDEN), (G_HIDDEN,), (G_HIDDEN,), (G_HIDDEN,),
            (G_HIDDEN, G_HIDDEN), (G_HIDDEN,), (G_HIDDEN, D), (D,)]


def _glorot(rng, shape):
    lim = np.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-lim, lim, size=shape)


def init_disc_params(D, rng, K=K_CLASSES):
    return [_glorot(rng, s) if len(s) == 2 else np.zeros(s) for s in disc_shapes(D, K)]


def init_gen_params(D, rng):
    p = [_glorot(rng, s) if len(s) == 2 else np.zeros(s) for s in gen_shapes(D)]
    p[2] = np.ones(G_HIDDEN)   # gamma
    return p


def flatten(params):
    return np.concatenate([np.asarray(p, dtype=np.float64).ravel() for p in params])


def unflatten(flat, shapes):
    out, o = [], 0
    for s in shapes:
        n = int(np.prod(s))
        out.append(np.array(flat[o:o + n], dtype=np.float64).reshape(s))
        o += n
    assert o == len(flat)
    return out


# ------------------------------------------------------------------ generator
def gen_forward(pG, z):
    """mr_gan.py:110-114: Dense500 softplus -> BN(batch stats) -> Dense500 softplus -> Dense D."""
    W1, b1, gamma, beta, W2, b2, W3, b3 = pG
    h1 = softplus(z @ W1 + b1)
    mu = h1.mean(axis=0)
    var = ((h1 - mu) ** 2).mean(axis=0)           # biased
    istd = 1.0 / np.sqrt(var + BN_EPS)
    xhat = (h1 - mu) * istd
    u = gamma * xhat + beta
    h2 = softplus(u @ W2 + b2)
    out = h2 @ W3 + b3
    return out, dict(z=z, h1=h1, istd=istd, xhat=xhat, u=u, h2=h2)


def gen_backward(pG, c, dout):
    W1, b1, gamma, beta, W2, b2, W3, b3 = pG
    B = dout.shape[0]
    gW3 = c['h2'].T @ dout
    gb3 = dout.sum(axis=0)
    dz2 = (dout @ W3.T) * (1.0 - np.exp(-c['h2']))     # softplus' = sigmoid(pre) = 1-exp(-softplus)
    gW2 = c['u'].T @ dz2
    gb2 = dz2.sum(axis=0)
    du = dz2 @ W2.T
    ggamma = (du * c['xhat']).sum(axis=0)
    gbeta = du.sum(axis=0)
    dxh = du * gamma
    dh1 = c['istd'] / B * (B * dxh - dxh.sum(axis=0) - c['xhat'] * (dxh * c['xhat']).sum(axis=0))
    dz1 = dh1 * (1.0 - np.exp(-c['h1']))
    gW1 = c['z'].T @ dz1
    gb1 = dz1.sum(axis=0)
    return [gW1, gb1, ggamma, gbeta, gW2, gb2, gW3, gb3]


# -------------------------------------------------------------- discriminator
# Variant of others/wganlpctsemi.py:166-179 (LeakyReLU + Dropout stacks), off by default:
#   alpha > 0 : the hidden activation is LeakyReLU(alpha) (Keras default alpha = 0.3) instead of ReLU;
#   dropout   : ``noise[l]`` for l >= 1 then holds Dropout KEEP FACTORS (0 or 1 / (1 - rate), Keras inverted dropout) that
#               MULTIPLY the layer input in place of the additive GaussianNoise; noise[0] stays the N(0,1) input noise.
def disc_forward(pD, x, noise=None, upto_mid=False, alpha=0.0, dropout=False):
    """mr_gan.py:117-128.  ``noise`` = list of 5 N(0,1) arrays (train phase) or None (test phase)."""
    a, a_in, hs = x, [], []
    for l in range(5):
        if noise is None:
            ai = a
        elif dropout and l >= 1:
            ai = a * noise[l]
        else:
            ai = a + D_SIGMAS[l] * noise[l]
        z = ai @ pD[2 * l] + pD[2 * l + 1]
        h = np.where(z > 0, z, alpha * z)
        a_in.append(ai)
        hs.append(h)
        a = h
    cache = dict(a_in=a_in, h=hs, alpha=alpha, drop=noise if (dropout and noise is not None) else None)
    if upto_mid:                       # mid_output model, mr_gan.py:127,133
        return a, cache
    return a @ pD[10] + pD[11], cache


def disc_backward(pD, c, dtop, from_mid=False, need_dx=False):
    """Gradients wrt D params (and optionally wrt the input)."""
    g = [None] * 12
    if from_mid:
        dh = dtop
    else:
        g[10] = c['h'][4].T @ dtop
        g[11] = dtop.sum(axis=0)
        dh = dtop @ pD[10].T
    alpha, drop = c.get('alpha', 0.0), c.get('drop')
    for l in range(4, -1, -1):
        dz = dh * np.where(c['h'][l] > 0, 1.0, alpha)
        g[2 * l] = c['a_in'][l].T @ dz
        g[2 * l + 1] = dz.sum(axis=0)
        if l > 0 or need_dx:
            dh = dz @ pD[2 * l].T                  # gradient w.r.t. the layer INPUT a_in[l] ...
            if drop is not None and l >= 1:
                dh = dh * drop[l]                  # ... and through the Dropout in front of it, w.r.t. h[l-1]
    return g, (dh if need_dx else None)


# --------------------------------------------------------------------- losses
def disc_losses(l_lab, labels, l_unl, l_fake):
    """mr_gan.py:146-149,161 -> (loss_lab, loss_unl, train_err, dl_lab, dl_unl, dl_fake)."""
    B = l_lab.shape[0]
    z_lab, z_unl, z_fake = logsumexp(l_lab), logsumexp(l_unl), logsumexp(l_fake)
    loss_lab = -l_lab[np.arange(B), labels].mean() + z_lab.mean()
    loss_unl = -0.5 * z_unl.mean() + 0.5 * softplus(z_unl).mean() + 0.5 * softplus(z_fake).mean()
    train_err = float((l_lab.argmax(axis=1) != labels).mean())
    onehot = np.zeros_like(l_lab)
    onehot[np.arange(B), labels] = 1.0
    dl_lab = (softmax(l_lab) - onehot) / B
    dl_unl = UNLABELED_WEIGHT * 0.5 * (sigmoid(z_unl) - 1.0)[:, None] * softmax(l_unl) / l_unl.shape[0]
    dl_fake = UNLABELED_WEIGHT * 0.5 * sigmoid(z_fake)[:, None] * softmax(l_fake) / l_fake.shape[0]
    return loss_lab, loss_unl, train_err, dl_lab, dl_unl, dl_fake


def fm_loss(f_fake, f_real):
    """mr_gan.py:152-154 -> (loss_gen, d loss / d f_fake)."""
    diff = f_fake.mean(axis=0) - f_real.mean(axis=0)
    loss = (diff ** 2).mean()
    d = np.broadcast_to(2.0 * diff / (diff.size * f_fake.shape[0]), f_fake.shape).copy()
    return loss, d


# ----------------------------------------------------------------------- Adam
def adam_update(params, grads, ms, vs, t, lr, b1, b2, eps):
    """Keras-2.0.9 ``Adam.get_updates`` (see module docstring).  In place."""
    lr_t = lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    for p, g, m, v in zip(params, grads, ms, vs):
        m *= b1
        m += (1.0 - b1) * g
        v *= b2
        v += (1.0 - b2) * g * g
        p -= lr_t * m / (np.sqrt(v) + eps)


class GanOracle:
    """One fold's G + D + the shared Adam (mr_gan.py:109-171)."""

    def __init__(self, pD, pG, shared_t=True, lr=GAN_LR, b1=GAN_B1, b2=GAN_B2, eps=GAN_EPS, alpha=0.0, dropout=False):
        self.var = dict(alpha=alpha, dropout=dropout)      # discriminator variant (see disc_forward)
        self.pD = [np.array(p, dtype=np.float64) for p in pD]
        self.pG = [np.array(p, dtype=np.float64) for p in pG]
        self.mD = [np.zeros_like(p) for p in self.pD]
        self.vD = [np.zeros_like(p) for p in self.pD]
        self.mG = [np.zeros_like(p) for p in self.pG]
        self.vG = [np.zeros_like(p) for p in self.pG]
        self.shared_t = shared_t
        self.iterations = 0            # Keras `adam.iterations` (shared)
        self.it_D = 0
        self.it_G = 0
        self.hp = (lr, b1, b2, eps)

    def _t(self, which):
        if self.shared_t:
            self.iterations += 1
            return self.iterations
        if which == 'D':
            self.it_D += 1
            return self.it_D
        self.it_G += 1
        return self.it_G

    # train_batch_disc, mr_gan.py:169 (graph at :141-149,161,166)
    def disc_grads(self, x_lab, labels, x_unl, z, n_lab, n_unl, n_fake):
        fake, _ = gen_forward(self.pG, z)
        l_lab, c_lab = disc_forward(self.pD, x_lab, n_lab, **self.var)
        l_unl, c_unl = disc_forward(self.pD, x_unl, n_unl, **self.var)
        l_fake, c_fake = disc_forward(self.pD, fake, n_fake, **self.var)
        ll, lu, te, d_lab, d_unl, d_fake = disc_losses(l_lab, labels, l_unl, l_fake)
        g = [np.zeros_like(p) for p in self.pD]
        for c, d in ((c_lab, d_lab), (c_unl, d_unl), (c_fake, d_fake)):
            gi, _ = disc_backward(self.pD, c, d)
            for a, b in zip(g, gi):
                a += b
        return (ll, lu, te), g

    def disc_step(self, x_lab, labels, x_unl, z, n_lab, n_unl, n_fake):
        out, g = self.disc_grads(x_lab, labels, x_unl, z, n_lab, n_unl, n_fake)
        adam_update(self.pD, g, self.mD, self.vD, self._t('D'), *self.hp)
        return out

    # train_batch_gen, mr_gan.py:170 (graph at :152-154,167)
    def gen_grads(self, x_unl, z, n_fake, n_real):
        fake, cg = gen_forward(self.pG, z)
        f_fake, c_fake = disc_forward(self.pD, fake, n_fake, upto_mid=True, **self.var)
        f_real, _ = disc_forward(self.pD, x_unl, n_real, upto_mid=True, **self.var)
        loss, dmid = fm_loss(f_fake, f_real)
        _, dfake = disc_backward(self.pD, c_fake, dmid, from_mid=True, need_dx=True)
        return loss, gen_backward(self.pG, cg, dfake)

    def gen_step(self, x_unl, z, n_fake, n_real):
        loss, g = self.gen_grads(x_unl, z, n_fake, n_real)
        adam_update(self.pG, g, self.mG, self.vG, self._t('G'), *self.hp)
        return loss

    # test_batch, mr_gan.py:171 (phase 0: GaussianNoise is identity)
    def test_batch(self, x, y):
        logits, _ = disc_forward(self.pD, x, None, alpha=self.var['alpha'])
        return float((logits.argmax(axis=1) != y).mean())


# ------------------------------------------------------------------ mr_nn step
class NnOracle:
    """mr_nn.py:101-118: D architecture as a classifier, MSE vs one-hot, default Adam."""

    def __init__(self, pD, lr=NN_LR, b1=NN_B1, b2=NN_B2, eps=NN_EPS, alpha=0.0, dropout=False):
        self.var = dict(alpha=alpha, dropout=dropout)
        self.pD = [np.array(p, dtype=np.float64) for p in pD]
        self.m = [np.zeros_like(p) for p in self.pD]
        self.v = [np.zeros_like(p) for p in self.pD]
        self.iterations = 0
        self.hp = (lr, b1, b2, eps)

    def grads(self, x, labels, noise):
        logits, c = disc_forward(self.pD, x, noise, **self.var)
        onehot = np.zeros_like(logits)
        onehot[np.arange(len(labels)), labels] = 1.0
        diff = logits - onehot
        loss = (diff ** 2).mean(axis=1).mean()          # Keras 'mse': mean over last axis, then batch
        acc = float((logits.argmax(axis=1) == labels).mean())
        g, _ = disc_backward(self.pD, c, 2.0 * diff / diff.size)
        return (loss, acc), g

    def step(self, x, labels, noise):
        out, g = self.grads(x, labels, noise)
        self.iterations += 1
        adam_update(self.pD, g, self.m, self.v, self.iterations, *self.hp)
        return out

    def evaluate(self, x, y):
        logits, _ = disc_forward(self.pD, x, None, alpha=self.var['alpha'])
        onehot = np.zeros_like(logits)
        onehot[np.arange(len(y)), y] = 1.0
        return float(((logits - onehot) ** 2).mean()), float((logits.argmax(axis=1) == y).mean())
