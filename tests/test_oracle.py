"""CPU suite, part 1: the oracle itself (Philox KAT, finite differences, autograd twin,
committed golden vectors).  PARITY UNPINNED by the reference -- see oracle/gan_oracle.py."""
import os

import numpy as np
import pytest
import torch

from oracle import fold_loop, gan_oracle as O, make_golden, philox, torch_twin as T


def test_philox_random123_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(g) for g in got) == want


def test_normal_stream_moments_and_indexing():
    key = philox.fold_key(1, 0)
    n = philox.normal(key, 3, 0, 2000, 501)
    assert abs(n.mean()) < 5e-3 and abs(n.std() - 1) < 5e-3 and abs((n ** 4).mean() - 3) < 0.05
    # a sub-block addressed by row0 equals the slice of the big block
    sub = philox.normal(key, 3, 0, 10, 501, row0=50)
    np.testing.assert_array_equal(sub, n[50:60])
    assert not np.allclose(philox.normal(key, 4, 0, 8, 8), n[:8, :8])       # step decorrelates
    assert not np.allclose(philox.normal(key, 3, 1, 8, 8), n[:8, :8])       # tensor id decorrelates


def _toy(D=12, B=6, seed=0):
    rng = np.random.default_rng(seed)
    pD, pG = O.init_disc_params(D, rng), O.init_gen_params(D, rng)
    for p in pD + pG:
        if p.ndim == 1:
            p += 0.1 * rng.standard_normal(p.shape)
    ws = (D,) + O.D_WIDTHS[:4]
    nz = lambda: [rng.standard_normal((B, w)) for w in ws]
    return dict(pD=pD, pG=pG, xl=rng.standard_normal((B, D)), xu=rng.standard_normal((B, D)),
                z=rng.standard_normal((B, O.NOISE_SIZE)), lab=rng.integers(0, 6, B), n=[nz(), nz(), nz()], rng=rng)


def test_disc_gradients_finite_differences():
    c = _toy()
    m = O.GanOracle(c['pD'], c['pG'])
    (_, _, _), g = m.disc_grads(c['xl'], c['lab'], c['xu'], c['z'], *c['n'])
    rng = np.random.default_rng(1)
    for ti in (0, 1, 4, 9, 10, 11):
        p = m.pD[ti]
        for _ in range(4):
            idx = tuple(rng.integers(0, s) for s in p.shape)
            old, eps = p[idx], 1e-5
            f = []
            for d in (eps, -eps):
                p[idx] = old + d
                (ll, lu, _), _g = m.disc_grads(c['xl'], c['lab'], c['xu'], c['z'], *c['n'])
                f.append(ll + lu)
            p[idx] = old
            fd = (f[0] - f[1]) / (2 * eps)
            assert abs(fd - g[ti][idx]) < 1e-6 + 1e-4 * abs(fd), (ti, idx, fd, g[ti][idx])


def test_gen_gradients_finite_differences():
    c = _toy()
    m = O.GanOracle(c['pD'], c['pG'])
    _, g = m.gen_grads(c['xu'], c['z'], c['n'][0], c['n'][1])
    rng = np.random.default_rng(2)
    for ti in range(8):
        p = m.pG[ti]
        for _ in range(4):
            idx = tuple(rng.integers(0, s) for s in p.shape)
            old, eps = p[idx], 1e-5
            f = []
            for d in (eps, -eps):
                p[idx] = old + d
                f.append(m.gen_grads(c['xu'], c['z'], c['n'][0], c['n'][1])[0])
            p[idx] = old
            fd = (f[0] - f[1]) / (2 * eps)
            assert abs(fd - g[ti][idx]) < 1e-9 + 2e-4 * abs(fd), (ti, idx, fd, g[ti][idx])


def test_numpy_oracle_matches_autograd_twin_float64():
    c = _toy(D=20, B=10)
    m, t = O.GanOracle(c['pD'], c['pG']), T.TorchGan(c['pD'], c['pG'], dtype=torch.float64)
    for _ in range(3):
        a = m.disc_step(c['xl'], c['lab'], c['xu'], c['z'], *c['n'])
        b = t.disc_step(c['xl'], c['lab'], c['xu'], c['z'], *c['n'])
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-12)
        ga = m.gen_step(c['xu'], c['z'], c['n'][0], c['n'][1])
        gb = t.gen_step(c['xu'], c['z'], c['n'][0], c['n'][1])
        np.testing.assert_allclose(ga, gb, rtol=1e-9)
    for p, q in zip(m.pD + m.pG, t.pD + t.pG):
        np.testing.assert_allclose(p, q.detach().numpy(), rtol=1e-8, atol=1e-10)
    assert m.iterations == 6 and t.iterations == 6                    # shared counter: 3 D + 3 G steps


def test_nn_oracle_matches_autograd_twin():
    c = _toy(D=20, B=10)
    m, t = O.NnOracle(c['pD']), T.TorchNn(c['pD'], dtype=torch.float64)
    for _ in range(3):
        a = m.step(c['xl'], c['lab'], c['n'][0])
        b = t.step(c['xl'], c['lab'], c['n'][0])
        np.testing.assert_allclose(a, b, rtol=1e-10)
    for p, q in zip(m.pD, t.pD):
        np.testing.assert_allclose(p, q.detach().numpy(), rtol=1e-8, atol=1e-10)


def test_shared_adam_counter_semantics():
    """mr_gan.py:165-167: one Adam object -> D step k sees t=2k-1, G step k sees t=2k."""
    c = _toy()
    shared, own = O.GanOracle(c['pD'], c['pG'], shared_t=True), O.GanOracle(c['pD'], c['pG'], shared_t=False)
    for m in (shared, own):
        m.disc_step(c['xl'], c['lab'], c['xu'], c['z'], *c['n'])
        m.gen_step(c['xu'], c['z'], c['n'][0], c['n'][1])
    assert shared.iterations == 2 and own.it_D == 1 and own.it_G == 1
    # first D step identical (t=1 both); first G step differs (t=2 vs t=1)
    np.testing.assert_allclose(shared.pD[0], own.pD[0])
    assert not np.allclose(shared.pG[0], own.pG[0])
    # closed form of the first Adam step: p -= lr_t * g/(|g|+eps) with m=(1-b1)g, v=(1-b2)g^2
    p, g = np.array([1.0]), np.array([0.3])
    mm, vv = np.zeros(1), np.zeros(1)
    O.adam_update([p], [g], [mm], [vv], 1, 6e-4, 0.5, 0.999, 1e-8)
    lr_t = 6e-4 * np.sqrt(1 - 0.999) / (1 - 0.5)
    np.testing.assert_allclose(p, 1.0 - lr_t * 0.15 / (np.sqrt(0.001 * 0.09) + 1e-8))


def test_loss_definitions_against_direct_formulas():
    """mr_gan.py:146-149,152-154,161 restated literally."""
    rng = np.random.default_rng(5)
    l_lab, l_unl, l_fake = (rng.standard_normal((7, 6)) * 2 for _ in range(3))
    y = rng.integers(0, 6, 7)
    ll, lu, te, *_ = O.disc_losses(l_lab, y, l_unl, l_fake)
    lse = lambda x: np.log(np.exp(x).sum(axis=1))
    assert np.isclose(ll, -l_lab[np.arange(7), y].mean() + lse(l_lab).mean())
    assert np.isclose(lu, -0.5 * lse(l_unl).mean() + 0.5 * np.log1p(np.exp(lse(l_unl))).mean()
                      + 0.5 * np.log1p(np.exp(lse(l_fake))).mean())
    assert np.isclose(te, np.mean(l_lab.argmax(1) != y))
    f, r = rng.standard_normal((7, 250)), rng.standard_normal((7, 250))
    assert np.isclose(O.fm_loss(f, r)[0], np.mean((f.mean(0) - r.mean(0)) ** 2))


@pytest.mark.parametrize("name", ["gan_steps_D36", "gan_steps_D30_B8", "gan_steps_D1200"])
def test_oracle_reproduces_committed_golden_vectors(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    key = tuple(int(k) for k in g['key'])
    m, losses = make_golden.run_case(int(g['D']), int(g['B']), int(g['seed']), int(g['n_pairs']), key)
    np.testing.assert_allclose(losses, g['losses'], rtol=1e-12)
    np.testing.assert_allclose(O.flatten(m.pD)[g['idxD']], g['pD_final'], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(O.flatten(m.pG)[g['idxG']], g['pG_final'], rtol=1e-12, atol=1e-15)


def test_noise_and_adam_known_answers(golden_dir):
    g = np.load(os.path.join(golden_dir, "noise_block.npz"))
    key = tuple(int(k) for k in g['key'])
    np.testing.assert_array_equal(philox.normal(key, 9, 2, 10, 7, row0=50), g['block'])
    a = np.load(os.path.join(golden_dir, "adam_kat.npz"))
    P, mm, vv = [a['p0'].copy()], np.zeros(257), np.zeros(257)
    for t in (1, 2, 3):
        O.adam_update(P, [a['g'] * t], [mm], [vv], t, O.GAN_LR, O.GAN_B1, O.GAN_B2, O.GAN_EPS)
    np.testing.assert_allclose(P[0], a['p3'], rtol=1e-13)


def test_epoch_loop_restatement_shapes_and_counters():
    rng = np.random.default_rng(0)
    D, B, N = 16, 10, 60
    X, y = rng.standard_normal((N, D)), np.repeat(np.arange(6), 10)
    Xtr, Xte, ytr, yte, lab_rows, unl = fold_loop.prep_fold(X, X[:12], y, y[:12], 0.2, None, rng)
    assert len(lab_rows) == 12 and unl is None
    np.testing.assert_allclose(Xtr.mean(0), 0, atol=1e-12)
    il, iu, iu2 = fold_loop.epoch_indices(rng, N, lab_rows)
    assert il.shape == iu.shape == iu2.shape == (N,)
    assert set(il) <= set(lab_rows) and sorted(iu) == list(range(N))
    m = O.GanOracle(O.init_disc_params(D, rng), O.init_gen_params(D, rng))
    stats, step = fold_loop.train_epoch(m, Xtr, ytr, il, iu, iu2, philox.fold_key(0, 0), 0, B=B)
    assert stats.shape == (N // B, 4) and step == 2 * (N // B) and m.iterations == step
    assert np.isfinite(stats).all()
    assert 0.0 <= fold_loop.eval_batches(m, Xte, yte, B=6) <= 1.0


def test_device_permutation_restatement_is_a_permutation_and_tiles_like_the_reference():
    """oracle/fold_loop.py:device_epoch_indices (what mrgan_train_epoch_seeded draws): every tile a permutation, the
    labeled stream tiled as mr_gan.py:189 (floor(N/L) permutations + a permutation of the first N mod L rows),
    deterministic in (key, epoch), different across epochs and streams."""
    from oracle import fold_loop, philox
    key = philox.fold_key(3, 4)
    lab = np.array([5, 9, 11, 40, 41, 77, 300])
    a = fold_loop.device_epoch_indices(key, 7, 31, lab)
    b = fold_loop.device_epoch_indices(key, 7, 31, lab)
    c = fold_loop.device_epoch_indices(key, 8, 31, lab)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)
    assert any(not np.array_equal(x, y) for x, y in zip(a, c))
    for j in range(4):
        assert sorted(a[0][7 * j:7 * j + 7]) == sorted(lab)
    assert sorted(a[0][28:]) == sorted(lab[:3])
    assert sorted(a[1]) == list(range(31)) and sorted(a[2]) == list(range(31)) and not np.array_equal(a[1], a[2])
    unl = np.arange(3, 15)
    u = fold_loop.device_epoch_indices(key, 1, 31, lab, unl)
    assert sorted(u[1][:12]) == list(unl) and sorted(u[2][24:]) == list(unl[:7])


def test_leaky_relu_dropout_variant_gradients_by_finite_differences():
    """The oracle's restatement of the LeakyReLU + Dropout discriminator stack (others/wganlpctsemi.py:166-179): analytic
    gradients of both steps against central differences; the replayed keep factors are 0 or 1 / (1 - rate) at about `rate`."""
    from oracle import fold_loop, gan_oracle as O, philox
    rng = np.random.default_rng(0)
    D, B, rate = 12, 8, 0.25
    pD, pG = O.init_disc_params(D, rng), O.init_gen_params(D, rng)
    for p in pD + pG:
        if p.ndim == 1:
            p += 0.1 * rng.standard_normal(p.shape)
    key = philox.fold_key(1, 2)
    m = O.GanOracle(pD, pG, alpha=0.3, dropout=True)
    x1, x2 = rng.standard_normal((B, D)), rng.standard_normal((B, D))
    y, z = rng.integers(0, 6, B), rng.standard_normal((B, 100))
    n = [fold_loop.d_transforms(key, 0, B, D, r0, rate) for r0 in (0, B, 2 * B)]
    f = np.concatenate([a.ravel() for a in n[0][1:]])
    assert set(np.unique(f)) <= {0.0, np.float64(np.float32(1) / (np.float32(1) - np.float32(rate)))} and 0.1 < (f == 0).mean() < 0.4
    (_, _, _), g = m.disc_grads(x1, y, x2, z, *n)
    for t in (0, 1, 4, 8, 10):
        idx = tuple(int(rng.integers(0, s)) for s in m.pD[t].shape)
        old, e = m.pD[t][idx], 1e-6
        vals = []
        for d in (e, -e):
            m.pD[t][idx] = old + d
            (a, b, _), _ = m.disc_grads(x1, y, x2, z, *n)
            vals.append(a + b)
        m.pD[t][idx] = old
        assert abs(g[t][idx] - (vals[0] - vals[1]) / (2 * e)) <= 1e-6 * max(1.0, abs(g[t][idx]) * 1e3)
    _, gg = m.gen_grads(x2, z, n[0], n[1])
    for t in (0, 2, 4, 6):
        idx = tuple(int(rng.integers(0, s)) for s in m.pG[t].shape)
        old, e = m.pG[t][idx], 1e-6
        vals = []
        for d in (e, -e):
            m.pG[t][idx] = old + d
            vals.append(m.gen_grads(x2, z, n[0], n[1])[0])
        m.pG[t][idx] = old
        assert abs(gg[t][idx] - (vals[0] - vals[1]) / (2 * e)) <= 1e-9 + 1e-4 * abs(gg[t][idx])
