"""Caffe op semantics of the reference (torch) engine: the oracle the CUDA kernels are tested against."""
import math

import pytest
import torch

from poseidon_b200.ops import reference as R


def test_pool_output_size_rule():
    assert R.pool_out_size(55, 3, 2, 0) == 27 and R.pool_out_size(13, 3, 2, 0) == 6
    assert R.pool_out_size(12, 3, 2, 0) == 6            # ceil (floor would give 5)
    assert R.pool_out_size(28, 3, 1, 1) == 28
    assert R.pool_out_size(4, 2, 2, 1) == 3             # window starting in the padding is dropped
    x = torch.randn(1, 2, 12, 12)
    assert R.max_pool(x, (3, 3), (2, 2), (0, 0)).shape[-1] == 6
    assert R.ave_pool(x, (3, 3), (2, 2), (0, 0)).shape[-1] == 6


def test_ave_pool_divisor_counts_padding_not_overhang():
    x = torch.ones(1, 1, 4, 4)
    y = R.ave_pool(x, (3, 3), (2, 2), (1, 1))
    # window at (0,0): rows -1..1, cols -1..1 -> 4 real ones / 9
    assert y[0, 0, 0, 0].item() == pytest.approx(4 / 9)
    y2 = R.ave_pool(torch.ones(1, 1, 5, 5), (3, 3), (2, 2), (0, 0))
    assert y2.shape[-1] == 2 and y2[0, 0, 1, 1].item() == pytest.approx(1.0)


def test_max_pool_mask_is_plane_index():
    x = torch.arange(16.0).reshape(1, 1, 4, 4)
    y, m = R.max_pool(x, (2, 2), (2, 2), (0, 0), return_mask=True)
    assert y.reshape(-1).tolist() == [5, 7, 13, 15] and m.reshape(-1).tolist() == [5, 7, 13, 15]


def test_lrn_across_matches_formula():
    x = torch.randn(2, 7, 3, 3)
    n, alpha, beta = 5, 0.3, 0.75
    y = R.lrn_across(x, n, alpha, beta)
    ref = torch.empty_like(x)
    for c in range(7):
        lo, hi = max(0, c - 2), min(7, c + 3)
        scale = 1 + (alpha / n) * (x[:, lo:hi] ** 2).sum(1)
        ref[:, c] = x[:, c] * scale ** (-beta)
    assert torch.allclose(y, ref, atol=1e-6)


def test_lrn_within_channel():
    x = torch.rand(1, 2, 5, 5)
    y = R.lrn_within(x, 3, 0.5, 0.75)
    s = R.ave_pool(x * x, (3, 3), (1, 1), (1, 1))[..., :5, :5]
    assert torch.allclose(y, x * (1 + 0.5 * s) ** -0.75, atol=1e-6)


def test_softmax_loss_and_gradient():
    x = torch.randn(6, 5, requires_grad=True)
    label = torch.tensor([0., 1, 2, 3, 4, 0])
    loss = R.softmax_loss(x, label)
    ref = torch.nn.functional.cross_entropy(x, label.long())
    assert loss.item() == pytest.approx(ref.item(), rel=1e-6)
    loss.backward()
    p = torch.softmax(x.detach(), 1)
    p[torch.arange(6), label.long()] -= 1
    assert torch.allclose(x.grad, p / 6, atol=1e-6)
    # spatial variant: (N, C, H, W) with per-pixel labels, normalised by N*H*W
    xs = torch.randn(2, 3, 2, 2)
    ls = torch.randint(0, 3, (2, 1, 2, 2)).float()
    l2 = R.softmax_loss(xs, ls)
    ref2 = torch.nn.functional.cross_entropy(xs, ls.long().squeeze(1))
    assert l2.item() == pytest.approx(ref2.item(), rel=1e-6)


def test_other_losses():
    a, b = torch.randn(4, 3), torch.randn(4, 3)
    assert R.euclidean_loss(a, b).item() == pytest.approx(((a - b) ** 2).sum().item() / 8, rel=1e-6)
    x = torch.randn(4, 3)
    lab = torch.tensor([0., 2, 1, 1])
    sign = torch.ones(4, 3)
    sign[torch.arange(4), lab.long()] = -1
    m = (1 + sign * x).clamp_min(0)
    assert R.hinge_loss(x, lab, "L1").item() == pytest.approx(m.sum().item() / 4, rel=1e-6)
    assert R.hinge_loss(x, lab, "L2").item() == pytest.approx((m * m).sum().item() / 4, rel=1e-6)
    t = torch.rand(4, 3)
    ref = torch.nn.functional.binary_cross_entropy_with_logits(x, t, reduction="sum") / 4
    assert R.sigmoid_cross_entropy_loss(x, t).item() == pytest.approx(ref.item(), rel=1e-5)
    p = torch.softmax(x, 1)
    assert R.multinomial_logistic_loss(p, lab).item() == pytest.approx(
        torch.nn.functional.nll_loss(p.log(), lab.long()).item(), rel=1e-6)
    H = torch.eye(3)
    assert R.infogain_loss(p, lab, H).item() == pytest.approx(R.multinomial_logistic_loss(p, lab).item(), rel=1e-6)
    sim = torch.tensor([1., 0, 1, 0])
    d2 = ((a - b) ** 2).sum(1)
    ref = (sim * d2 + (1 - sim) * (1.0 - d2).clamp_min(0)).sum() / 8
    assert R.contrastive_loss(a, b, sim, 1.0).item() == pytest.approx(ref.item(), rel=1e-6)


def test_neurons_and_misc():
    x = torch.tensor([-2.0, -0.5, 0.0, 0.5, 2.0])
    assert R.relu(x, 0.1).tolist() == pytest.approx([-0.2, -0.05, 0, 0.5, 2.0])
    assert torch.allclose(R.bnll(x), torch.log1p(torch.exp(x)), atol=1e-6)
    assert torch.allclose(R.power(x, 2, 0.5, 1.0), (1 + 0.5 * x) ** 2)
    v = torch.randn(2, 3, 4, 4)
    y = R.mvn(v)
    assert y.reshape(6, -1).mean(1).abs().max() < 1e-5 and (y.reshape(6, -1).std(1, unbiased=False) - 1).abs().max() < 1e-3
    acc = R.accuracy(torch.tensor([[0.1, 0.9], [0.8, 0.2]]), torch.tensor([1., 1.]))
    assert acc.item() == 0.5
    am = R.argmax(torch.tensor([[0.1, 0.9, 0.5]]), top_k=2, out_max_val=True)
    assert am.shape == (1, 2, 2, 1) and am[0, 0, :, 0].tolist() == [1, 2]
    e = R.eltwise([torch.ones(2), 2 * torch.ones(2)], "SUM", [1.0, -1.0])
    assert e.tolist() == [-1, -1]
    assert R.eltwise([torch.tensor([1., 5]), torch.tensor([3., 2])], "MAX").tolist() == [3, 5]


def test_dropout_train_scale_and_test_identity():
    x = torch.ones(10000)
    y = R.dropout(x, 0.25, True)
    assert abs((y > 0).float().mean().item() - 0.75) < 0.03
    assert y.max().item() == pytest.approx(1 / 0.75)
    assert R.dropout(x, 0.25, False) is x


def test_stochastic_pool_test_mode():
    x = torch.rand(1, 1, 4, 4) + 0.1
    y = R.stochastic_pool(x, (2, 2), (2, 2), train=False)
    w = x[0, 0, :2, :2]
    assert y[0, 0, 0, 0].item() == pytest.approx(((w * w).sum() / w.sum()).item(), rel=1e-5)
    yt = R.stochastic_pool(x, (2, 2), (2, 2), train=True)
    assert yt[0, 0, 0, 0].item() in [pytest.approx(v.item()) for v in w.reshape(-1)]


def test_optimizer_rules():
    w0, g = torch.tensor([1.0, -2.0]), torch.tensor([0.5, 0.25])
    w, h = w0.clone(), torch.tensor([0.1, 0.1])
    R.sgd_step(w, g, h, 0.1, 0.9, 0.01)
    hh = 0.1 * (g + 0.01 * w0) + 0.9 * 0.1
    assert torch.allclose(h, hh) and torch.allclose(w, w0 - hh)
    w, h = w0.clone(), torch.tensor([0.1, 0.1])
    R.nesterov_step(w, g, h, 0.1, 0.9, 0.0)
    hn = 0.1 * g + 0.09
    assert torch.allclose(w, w0 - (1.9 * hn - 0.9 * 0.1))
    w, h = w0.clone(), torch.zeros(2)
    R.adagrad_step(w, g, h, 0.1, 1e-8, 0.0)
    assert torch.allclose(h, g * g) and torch.allclose(w, w0 - 0.1 * g / (g.abs() + 1e-8))
    w, h = w0.clone(), torch.zeros(2)
    R.sgd_step(w, g, h, 0.1, 0.0, 0.01, l1=True)
    assert torch.allclose(w, w0 - 0.1 * (g + 0.01 * torch.sign(w0)))


def test_fillers():
    from poseidon_b200 import proto as P
    from poseidon_b200.layers import fill, set_filler_seed
    set_filler_seed(3)
    # xavier / positive_unitball follow the reference's 4-D blob arithmetic (fan_in = count / num): a convolution weight
    # (Cout, Cin, kh, kw) has num = Cout, an inner-product weight is the blob (1, 1, N, K) with num = 1
    c = torch.empty(64, 5, 3, 3)
    fill(c, P.FillerParameter(type="xavier"))
    assert c.abs().max().item() <= math.sqrt(3 / 45) + 1e-6 and c.std().item() > 0.1
    fill(c, P.FillerParameter(type="positive_unitball"))
    assert torch.allclose(c.view(64, -1).sum(1), torch.ones(64), atol=1e-5)
    t = torch.empty(64, 50)
    fill(t, P.FillerParameter(type="xavier"))
    assert t.abs().max().item() <= math.sqrt(3 / (64 * 50)) + 1e-6 and t.std().item() > 0.01
    fill(t, P.FillerParameter(type="gaussian", mean=1.0, std=0.5))
    assert abs(t.mean().item() - 1) < 0.05 and abs(t.std().item() - 0.5) < 0.05
    fill(t, P.FillerParameter(type="uniform", min=-2, max=-1))
    assert -2 <= t.min().item() and t.max().item() <= -1
    fill(t, P.FillerParameter(type="positive_unitball"))
    assert abs(t.sum().item() - 1.0) < 1e-4
    fill(t, P.FillerParameter(type="constant", value=0.25))
    assert torch.all(t == 0.25)
    fill(t, P.FillerParameter(type="gaussian", std=1.0, sparse=8))
    assert abs((t != 0).float().mean().item() - 8 / 64) < 0.03
    set_filler_seed(None)
