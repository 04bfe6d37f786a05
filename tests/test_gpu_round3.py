"""GPU tests added in round 3 (pytest -m gpu): the inference stem in one kernel (conv1 + bn1 statistics + max-pool of the raw
output in the epilogue) against the three-kernel stem it replaces - bit for bit - and against torch's max_pool2d."""
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))


@pytest.fixture(scope="module")
def bu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import gpu_bringup
    return gpu_bringup


# (B, G, input channels): G = 5 / 9 leave a partial block of 4 samples (9: a column set without any sample), B = 18 with
# G = 30 takes 16 pooled rows per unit (the cfg2 plan), the small batches 2 rows per unit (halo row on every unit)
@pytest.mark.parametrize("B,G,cin", [(3, 5, 3), (2, 3, 1), (5, 9, 3), (18, 30, 3), (2, 16, 1)])
def test_stem_conv_pool_is_bit_identical_to_conv_then_bn_relu_maxpool(bu, B, G, cin):
    """mauv_stem_conv_pool_f16 + bn_finalize + bn_act == mauv_gemm_f16 (stacked) + bn_finalize + mauv_bn_relu_maxpool_f16 with
    torch.equal on the statistics and on the activations (reference: conv1 -> bn1 -> relu -> maxpool of torchvision's
    ResNet._forward_impl, called from models/base_models.py:74-76); some BatchNorm weights are negative (window min)."""
    from mauv import ops
    torch.manual_seed(B * 100 + G)
    x = torch.randn(B, cin, 256, 256, device="cuda")
    a0 = ops.stem_im2col_f16(x, 7, 7, 2, 3)
    Kp = a0.shape[1]
    w = (torch.randn(G, 64, Kp, device="cuda") * 0.05).half()
    w[:, :, 49 * cin:] = 0                                    # the padding columns of the K order carry no weight
    gamma = torch.randn(64, device="cuda")
    gamma[::7] = 0.0
    beta = torch.randn(64, device="cuda") * 0.1
    Ho = Wo = 128
    # the path it replaces
    y, st = ops.gemm_f16(a0, w, stats=True, shared_a=True)
    ss = ops.bn_finalize(st, B * Ho * Wo, gamma, beta)
    ref = ops.bn_relu_maxpool_f16(y.view(G * B, Ho, Wo, 64), ss, G)
    # one kernel
    yp, st2 = ops.stem_conv_pool_f16(a0, w, B, Ho, gamma=gamma)
    assert st2.shape == st.shape and torch.equal(st2, st)
    ss2 = ops.bn_finalize(st2, B * Ho * Wo, gamma, beta)
    got = ops.bn_act_f16(yp, ss2, G, 64, relu=True)
    assert torch.equal(got, ref)
    # the pooled raw tensor itself: window max of y, window min where gamma < 0 (fp16 values, exact)
    sgn = torch.where(gamma < 0, -1.0, 1.0).view(1, 64, 1, 1)
    yn = y.view(G * B, Ho, Wo, 64).permute(0, 3, 1, 2).float()
    pooled = sgn * torch.nn.functional.max_pool2d(yn * sgn, 3, 2, 1)
    assert torch.equal(yp.permute(0, 3, 1, 2).float(), pooled)
    # gamma = None means "all scales non-negative"
    yq, _ = ops.stem_conv_pool_f16(a0, w, B, Ho, gamma=None)
    assert torch.equal(yq.permute(0, 3, 1, 2).float(), torch.nn.functional.max_pool2d(yn, 3, 2, 1))


@pytest.mark.parametrize("kind", ["multimodal", "unimodal"])
def test_engine_with_the_fused_stem_gives_bit_identical_logits(bu, kind):
    """MCEngine at 256 x 256 (the size of every BASELINE config) with and without the fused stem: torch.equal logits, for
    a sample group that fills the 4-sample blocks unevenly, and BatchNorm running statistics updated identically."""
    import bnn_oracle as O
    from mauv.engine import MCEngine
    _, model = bu.build_pair(kind)
    img, bathy, sss, _ = O.synthetic_batch(2, size=256)
    xs = [t.cuda() for t in ((img, bathy, sss) if kind == "multimodal" else (img,))]
    stem_bn = model.image_model_feat.bn1 if kind == "multimodal" else model.model.bn1
    with torch.no_grad():
        stem_bn.weight[3] = -0.7
        stem_bn.weight[10] = 0.0
    eng = MCEngine(model)
    assert eng.stem_pool
    rm0 = stem_bn.running_mean.clone()
    fused = eng.forward_mc(xs, 6, seed=77, group=6, sample0=0)
    rm_fused = stem_bn.running_mean.clone()
    stem_bn.running_mean.copy_(rm0)
    eng.stem_pool = False
    plain = eng.forward_mc(xs, 6, seed=77, group=6, sample0=0)
    assert torch.equal(fused, plain)
    assert torch.equal(rm_fused, stem_bn.running_mean)
    eng.stem_pool = True
    assert torch.equal(eng.forward_mc(xs, 6, seed=77, group=4, sample0=0), plain)
