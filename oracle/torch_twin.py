"""ORACLE (test infrastructure only) -- torch-CPU autograd twin of oracle/gan_oracle.py.

PARITY UNPINNED (see gan_oracle.py header).  Two jobs:
  1. cross-check the hand-derived gradients of the numpy oracle with an
     independent autograd implementation (run in float64 for that);
  2. be the timed "restated CPU baseline" (float32, all host threads, Python
     loop with two separate step calls per iteration and host-side noise, like
     mr_gan.py:204-213) for bench.py's cpu_baseline and ``--impl reference``.
     It is NOT Keras 2.0.9 / Theano 0.9.0 -- those cannot be installed here.

Never imported by the product path.
"""
import math
import torch
import torch.nn.functional as F

from . import gan_oracle as O


class _Adam:
    """Keras-2.0.9 Adam on a list of tensors (see gan_oracle.adam_update)."""

    def __init__(self, params, lr, b1, b2, eps):
        self.params = params
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps

    @torch.no_grad()
    def apply(self, grads, t):
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
        torch._foreach_mul_(self.m, self.b1)
        torch._foreach_add_(self.m, grads, alpha=1.0 - self.b1)
        torch._foreach_mul_(self.v, self.b2)
        torch._foreach_addcmul_(self.v, grads, grads, value=1.0 - self.b2)
        den = torch._foreach_sqrt(self.v)
        torch._foreach_add_(den, self.eps)
        torch._foreach_addcdiv_(self.params, self.m, den, value=-lr_t)


def _disc(pD, x, noise, upto_mid=False):
    a = x
    for l in range(5):
        if noise is not None:
            a = a + O.D_SIGMAS[l] * noise[l]
        a = torch.relu(a @ pD[2 * l] + pD[2 * l + 1])
    if upto_mid:
        return a
    return a @ pD[10] + pD[11]


def _gen(pG, z):
    W1, b1, gamma, beta, W2, b2, W3, b3 = pG
    h1 = F.softplus(z @ W1 + b1)
    mu = h1.mean(dim=0)
    var = ((h1 - mu) ** 2).mean(dim=0)
    u = gamma * (h1 - mu) / torch.sqrt(var + O.BN_EPS) + beta
    h2 = F.softplus(u @ W2 + b2)
    return h2 @ W3 + b3


class TorchGan:
    def __init__(self, pD, pG, dtype=torch.float32, shared_t=True):
        cv = lambda p: torch.tensor(p, dtype=dtype).requires_grad_(True)
        self.dtype = dtype
        self.pD = [cv(p) for p in pD]
        self.pG = [cv(p) for p in pG]
        self.adamD = _Adam(self.pD, O.GAN_LR, O.GAN_B1, O.GAN_B2, O.GAN_EPS)
        self.adamG = _Adam(self.pG, O.GAN_LR, O.GAN_B1, O.GAN_B2, O.GAN_EPS)
        self.shared_t = shared_t
        self.iterations = 0
        self.it = {'D': 0, 'G': 0}

    def _t(self, which):
        if self.shared_t:
            self.iterations += 1
            return self.iterations
        self.it[which] += 1
        return self.it[which]

    def _noise(self, rows, D):
        widths = (D,) + O.D_WIDTHS[:4]
        return [torch.randn(rows, w, dtype=self.dtype) for w in widths]

    def _cv(self, a):
        return torch.as_tensor(a, dtype=self.dtype)

    def disc_loss(self, x_lab, labels, x_unl, z, n_lab=None, n_unl=None, n_fake=None):
        x_lab, x_unl, z = self._cv(x_lab), self._cv(x_unl), self._cv(z)
        labels = torch.as_tensor(labels, dtype=torch.long)
        B, D = x_lab.shape
        cvn = lambda n, rows: self._noise(rows, D) if n is None else [self._cv(a) for a in n]
        n_lab, n_unl, n_fake = cvn(n_lab, B), cvn(n_unl, x_unl.shape[0]), cvn(n_fake, z.shape[0])
        fake = _gen(self.pG, z)
        l_lab = _disc(self.pD, x_lab, n_lab)
        l_unl = _disc(self.pD, x_unl, n_unl)
        l_fake = _disc(self.pD, fake, n_fake)
        z_unl = torch.logsumexp(l_unl, dim=1)
        loss_lab = -l_lab[torch.arange(B), labels].mean() + torch.logsumexp(l_lab, dim=1).mean()
        loss_unl = (-0.5 * z_unl.mean() + 0.5 * F.softplus(z_unl).mean()
                    + 0.5 * F.softplus(torch.logsumexp(l_fake, dim=1)).mean())
        err = (l_lab.argmax(dim=1) != labels).double().mean()
        return loss_lab, loss_unl, err

    def disc_step(self, *a, **k):
        ll, lu, err = self.disc_loss(*a, **k)
        grads = torch.autograd.grad(ll + O.UNLABELED_WEIGHT * lu, self.pD)
        self.adamD.apply(list(grads), self._t('D'))
        return float(ll.detach()), float(lu.detach()), float(err)

    def gen_loss(self, x_unl, z, n_fake=None, n_real=None):
        x_unl, z = self._cv(x_unl), self._cv(z)
        D = x_unl.shape[1]
        cvn = lambda n, rows: self._noise(rows, D) if n is None else [self._cv(a) for a in n]
        n_fake, n_real = cvn(n_fake, z.shape[0]), cvn(n_real, x_unl.shape[0])
        f_fake = _disc(self.pD, _gen(self.pG, z), n_fake, upto_mid=True)
        f_real = _disc(self.pD, x_unl, n_real, upto_mid=True)
        return ((f_fake.mean(dim=0) - f_real.mean(dim=0)) ** 2).mean()

    def gen_step(self, *a, **k):
        loss = self.gen_loss(*a, **k)
        grads = torch.autograd.grad(loss, self.pG)
        self.adamG.apply(list(grads), self._t('G'))
        return float(loss.detach())

    @torch.no_grad()
    def test_batch(self, x, y):
        logits = _disc(self.pD, self._cv(x), None)
        return float((logits.argmax(dim=1) != torch.as_tensor(y, dtype=torch.long)).double().mean())


class TorchNn:
    def __init__(self, pD, dtype=torch.float32):
        self.dtype = dtype
        self.pD = [torch.tensor(p, dtype=dtype).requires_grad_(True) for p in pD]
        self.adam = _Adam(self.pD, O.NN_LR, O.NN_B1, O.NN_B2, O.NN_EPS)
        self.iterations = 0

    def step(self, x, labels, noise=None):
        x = torch.as_tensor(x, dtype=self.dtype)
        labels = torch.as_tensor(labels, dtype=torch.long)
        if noise is None:
            widths = (x.shape[1],) + O.D_WIDTHS[:4]
            noise = [torch.randn(x.shape[0], w, dtype=self.dtype) for w in widths]
        else:
            noise = [torch.as_tensor(n, dtype=self.dtype) for n in noise]
        logits = _disc(self.pD, x, noise)
        loss = ((logits - F.one_hot(labels, O.K_CLASSES).to(self.dtype)) ** 2).mean(dim=1).mean()
        grads = torch.autograd.grad(loss, self.pD)
        self.iterations += 1
        self.adam.apply(list(grads), self.iterations)
        return float(loss.detach()), float((logits.argmax(dim=1) == labels).double().mean())
