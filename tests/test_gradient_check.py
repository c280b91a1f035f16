"""Caffe's GradientChecker, re-created: every differentiable layer type is checked against central finite differences in
float64 on CPU — w.r.t. all bottoms and all learnable blobs (upstream: src/caffe/test/test_gradient_check_util.hpp, typed
tests per layer; that directory is absent from the reference snapshot)."""
import pytest
import torch

from poseidon_b200 import proto as P
from poseidon_b200.layers import NetContext, create_layer
from poseidon_b200.layers.base import set_filler_seed
from poseidon_b200.proto import parse_text

CASES = [
    # (id, layer prototxt, bottom shapes, dict(non_diff_bottoms=set(), positive=False))
    ("conv_basic", 'type: CONVOLUTION convolution_param { num_output: 4 kernel_size: 3 stride: 1 pad: 1 '
                   'weight_filler { type: "gaussian" std: 0.5 } bias_filler { type: "gaussian" std: 0.5 } }', [(2, 3, 5, 6)], {}),
    ("conv_group_stride", 'type: CONVOLUTION convolution_param { num_output: 6 kernel_size: 3 stride: 2 pad: 1 group: 2 '
                          'weight_filler { type: "gaussian" std: 0.5 } bias_filler { type: "constant" value: 0.1 } }', [(2, 4, 7, 7)], {}),
    ("conv_rect_nobias", 'type: CONVOLUTION convolution_param { num_output: 3 kernel_h: 2 kernel_w: 3 stride_h: 1 stride_w: 2 '
                         'pad_h: 0 pad_w: 1 bias_term: false weight_filler { type: "gaussian" std: 0.5 } }', [(1, 2, 5, 6)], {}),
    ("pool_max", 'type: POOLING pooling_param { pool: MAX kernel_size: 3 stride: 2 }', [(2, 3, 7, 7)], {}),
    ("pool_max_pad", 'type: POOLING pooling_param { pool: MAX kernel_size: 3 stride: 2 pad: 1 }', [(2, 2, 6, 6)], {}),
    ("pool_ave_pad", 'type: POOLING pooling_param { pool: AVE kernel_size: 3 stride: 2 pad: 1 }', [(2, 2, 6, 7)], {}),
    ("lrn_across", 'type: LRN lrn_param { local_size: 5 alpha: 0.1 beta: 0.75 }', [(2, 7, 3, 3)], {}),
    ("lrn_within", 'type: LRN lrn_param { local_size: 3 alpha: 0.1 beta: 0.75 norm_region: WITHIN_CHANNEL }', [(2, 3, 5, 5)], {}),
    ("inner_product", 'type: INNER_PRODUCT inner_product_param { num_output: 5 weight_filler { type: "gaussian" std: 0.5 } '
                      'bias_filler { type: "gaussian" std: 0.5 } }', [(3, 2, 2, 3)], {}),
    ("relu", 'type: RELU', [(2, 3, 4, 4)], {}),
    ("relu_leaky", 'type: RELU relu_param { negative_slope: 0.1 }', [(2, 3, 4, 4)], {}),
    ("sigmoid", 'type: SIGMOID', [(2, 3, 2, 2)], {}),
    ("tanh", 'type: TANH', [(2, 3, 2, 2)], {}),
    ("absval", 'type: ABSVAL', [(2, 3, 2, 2)], {}),
    ("bnll", 'type: BNLL', [(2, 3, 2, 2)], {}),
    ("power", 'type: POWER power_param { power: 2.0 scale: 0.5 shift: 1.5 }', [(2, 3, 2, 2)], {}),
    ("softmax", 'type: SOFTMAX', [(3, 5, 2, 2)], {}),
    ("concat", 'type: CONCAT', [(2, 2, 3, 3), (2, 3, 3, 3)], {}),
    ("concat_num", 'type: CONCAT concat_param { concat_dim: 0 }', [(1, 2, 3, 3), (2, 2, 3, 3)], {}),
    ("slice", 'type: SLICE slice_param { slice_dim: 1 slice_point: 2 }', [(2, 5, 3, 3)], {"n_tops": 2}),
    ("split", 'type: SPLIT', [(2, 3, 2, 2)], {"n_tops": 2}),
    ("flatten", 'type: FLATTEN', [(2, 3, 2, 2)], {}),
    ("eltwise_sum", 'type: ELTWISE eltwise_param { operation: SUM coeff: 1.5 coeff: -0.5 }', [(2, 3, 2, 2), (2, 3, 2, 2)], {}),
    ("eltwise_prod", 'type: ELTWISE eltwise_param { operation: PROD }', [(2, 3, 2, 2), (2, 3, 2, 2)], {}),
    ("eltwise_max", 'type: ELTWISE eltwise_param { operation: MAX }', [(2, 3, 2, 2), (2, 3, 2, 2)], {}),
    ("mvn", 'type: MVN', [(2, 3, 4, 4)], {}),
    ("mvn_across_novar", 'type: MVN mvn_param { across_channels: true normalize_variance: false }', [(2, 3, 4, 4)], {}),
    ("dropout_test", 'type: DROPOUT dropout_param { dropout_ratio: 0.4 }', [(2, 3, 2, 2)], {"phase": "TEST"}),
    ("softmax_loss", 'type: SOFTMAX_LOSS', [(4, 5, 1, 1), "labels:5"], {}),
    ("softmax_loss_spatial", 'type: SOFTMAX_LOSS', [(2, 4, 2, 3), "labels_spatial:4"], {}),
    ("euclidean_loss", 'type: EUCLIDEAN_LOSS', [(3, 4, 1, 1), (3, 4, 1, 1)], {}),
    ("sigmoid_ce_loss", 'type: SIGMOID_CROSS_ENTROPY_LOSS', [(3, 4, 1, 1), "targets01"], {}),
    ("hinge_l1", 'type: HINGE_LOSS', [(4, 5, 1, 1), "labels:5"], {}),
    ("hinge_l2", 'type: HINGE_LOSS hinge_loss_param { norm: L2 }', [(4, 5, 1, 1), "labels:5"], {}),
    ("multinomial_logistic", 'type: MULTINOMIAL_LOGISTIC_LOSS', ["probs:5", "labels:5"], {}),
    ("contrastive", 'type: CONTRASTIVE_LOSS contrastive_loss_param { margin: 1.0 }', [(4, 3, 1, 1), (4, 3, 1, 1), "sim"], {}),
]


def _make_bottoms(specs, g):
    out, diff = [], []
    n = next((s[0] for s in specs if isinstance(s, tuple)), 4)
    spatial = next((s[2:] for s in specs if isinstance(s, tuple)), (1, 1))
    for s in specs:
        if isinstance(s, tuple):
            out.append(torch.randn(*s, dtype=torch.float64, generator=g))
            diff.append(True)
        elif s.startswith("labels_spatial:"):
            out.append(torch.randint(0, int(s.split(":")[1]), (n, 1) + tuple(spatial), generator=g).double())
            diff.append(False)
        elif s.startswith("labels:"):
            out.append(torch.randint(0, int(s.split(":")[1]), (n, 1, 1, 1), generator=g).double())
            diff.append(False)
        elif s == "targets01":
            out.append(torch.rand(*out[0].shape, dtype=torch.float64, generator=g))
            diff.append(False)
        elif s == "sim":
            out.append(torch.randint(0, 2, (n, 1, 1, 1), generator=g).double())
            diff.append(False)
        elif s.startswith("probs:"):
            k = int(s.split(":")[1])
            p = torch.rand(4, k, 1, 1, dtype=torch.float64, generator=g) + 0.1
            out.append(p / p.sum(1, keepdim=True))
            diff.append(True)
    return out, diff


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_layer_gradients_match_finite_differences(case):
    name, txt, specs, opts = case
    g = torch.Generator().manual_seed(1701)
    set_filler_seed(1701)
    lp = parse_text(f'name: "{name}" {txt}', P.LayerParameter)
    phase = P.TEST if opts.get("phase") == "TEST" else P.TRAIN
    layer = create_layer(lp, NetContext(phase=phase, dtype=torch.float64))
    bottoms, diffable = _make_bottoms(specs, g)
    if name == "multinomial_logistic":
        bottoms[1] = torch.randint(0, 5, (4, 1, 1, 1), generator=g).double()
    n_tops = opts.get("n_tops", 1)
    layer.check_blob_counts(len(bottoms), n_tops)
    layer.n_tops = n_tops                      # the Net assigns this from the prototxt's `top:` count
    layer.setup([tuple(b.shape) for b in bottoms])
    layer.double()
    pnames = [n for n, _ in layer.named_parameters()]
    params = [p.detach().clone().requires_grad_(True) for _, p in layer.named_parameters()]
    inputs = [b.clone().requires_grad_(d) for b, d in zip(bottoms, diffable)]

    def fn(*args):
        xs, ps = args[:len(inputs)], args[len(inputs):]
        out = torch.func.functional_call(layer, dict(zip(pnames, ps)), tuple(xs))
        outs = out if isinstance(out, (tuple, list)) else (out,)
        return tuple(o for o in outs if o.requires_grad or o.dtype.is_floating_point)

    assert torch.autograd.gradcheck(fn, tuple(inputs) + tuple(params), eps=1e-6, atol=1e-5, rtol=1e-4, nondet_tol=0.0,
                                    check_undefined_grad=False)
