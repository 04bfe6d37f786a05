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
    eng.stem_colsum = False      # (the by-product column sums differ from colsum_f16's in summation order only; checked below)
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
    eng.stem_colsum = True       # production: first moment of layer1.0's downsample statistics from the stem's bn_act pass
    got = eng.forward_mc(xs, 6, seed=77, group=6, sample0=0)
    assert (got - plain).abs().max().item() <= (1.5e-3 if kind == "multimodal" else 0.15) * plain.abs().max().item()


def _raw_and_bn(G, M, K, seed):
    torch.manual_seed(seed)
    y = torch.randn(G, M, K, device="cuda").half()
    ss = torch.stack([torch.randn(G, K, device="cuda") * 0.7, torch.randn(G, K, device="cuda") * 0.3], dim=-1).contiguous()
    return y, ss


@pytest.mark.parametrize("G,M,K", [(2, 8192, 64), (3, 4096 * 3, 128), (2, 4096, 256), (1, 64 * 5, 64)])
def test_second_moments_of_bn_relu_of_a_raw_tensor_without_materialising_it(bu, G, M, K):
    """mauv_gram_bn_f16 (BatchNorm + ReLU applied to the operand tiles in shared memory) against bn_act -> second moments of
    the materialised tensor: the a^T a partials must be bit-identical (same fp16 operands, same chunking), the column sums
    equal to fp32 summation-order noise, and the BatchNorm scale/shift that follow within 1e-6."""
    from mauv import _lib, ops
    lib = _lib.require_device()
    y, ss = _raw_and_bn(G, M, K, 11 + K)
    a, cs = ops.bn_act_f16(y, ss, G, K, relu=True, colsum=True)
    splits = ops.gram_splits(M, G, K)
    assert splits > 0
    a4 = a.view(G, M, 1, K)
    ref = ops.wgrad_f16(a4, a4, G, splits, 1, 1, 1, 0)
    gram = torch.empty_like(ref)
    colsum = torch.empty(G, splits, K, device="cuda")
    _lib.check(lib.mauv_gram_bn_f16(y.data_ptr(), ss.data_ptr(), gram.data_ptr(), colsum.data_ptr(), G, splits, M, K,
                                    torch.cuda.current_stream().cuda_stream))
    assert torch.equal(gram, ref)
    cs_ref = a.float().view(G, splits, M // splits, K).sum(2)
    assert torch.allclose(colsum, cs_ref, rtol=2e-6, atol=1e-3)
    assert torch.allclose(colsum.sum(1), cs.sum(1), rtol=1e-5, atol=1e-2)
    w = (torch.randn(G, 4 * K, K, device="cuda") * 0.05).half()
    gamma, beta = torch.rand(4 * K, device="cuda") + 0.5, torch.randn(4 * K, device="cuda") * 0.1
    s_ref = ops.bn_stats_from_gram(a, cs, w, M, gamma, beta, 1e-5, 0.1)
    s_xf = ops.bn_stats_from_gram(y, None, w, M, gamma, beta, 1e-5, 0.1, a_ss=ss)
    assert torch.allclose(s_xf, s_ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("G,M,N,K,res", [(2, 4096, 256, 64, True), (3, 1000, 512, 128, True), (2, 2048, 1024, 256, True),
                                         (2, 777, 256, 64, False)])
def test_fused_tail_on_the_raw_conv2_output_is_bit_identical_to_bn_act_then_fused_tail(bu, G, M, N, K, res):
    """mauv_gemm_bn_xf_f16 == mauv_bn_act_f16 -> mauv_gemm_bn_f16 (mode 2), torch.equal: the transform warps produce exactly
    bn_act's fp16 values, so the tensor core sees the same operands (ragged M: the zero-filled rows are never stored)."""
    from mauv import ops
    y, ss = _raw_and_bn(G, M, K, 5 + N)
    w = (torch.randn(G, N, K, device="cuda") * 0.05).half()
    ss3 = torch.stack([torch.rand(G, N, device="cuda") + 0.5, torch.randn(G, N, device="cuda") * 0.2], dim=-1).contiguous()
    r = torch.randn(G, M, N, device="cuda").half() if res else None
    a = ops.bn_act_f16(y, ss, G, K, relu=True)
    ref = ops.gemm_bn_act_f16(a, w, ss3, residual=r, relu=True)
    got = ops.gemm_bn_act_f16(y, w, ss3, residual=r, relu=True, a_ss=ss)
    assert torch.equal(got, ref)


@pytest.mark.parametrize("G,M,N,K1,K2", [(2, 4096, 256, 64, 64), (2, 1500, 512, 128, 256), (1, 2048, 1024, 256, 512)])
def test_k_concatenated_tail_with_a_raw_first_operand(bu, G, M, N, K1, K2):
    """mauv_gemm_bn_cat_xf_f16: only the a2 k-blocks are transformed, the downsample branch's input is read as is."""
    from mauv import ops
    y, ss = _raw_and_bn(G, M, K1, 3 + N)
    x = torch.randn(G, M, K2, device="cuda").half()
    wcat = (torch.randn(G, N, K1 + K2, device="cuda") * 0.05).half()
    shift = torch.stack([torch.ones(G, N, device="cuda"), torch.randn(G, N, device="cuda") * 0.2], dim=-1).contiguous()
    a = ops.bn_act_f16(y, ss, G, K1, relu=True)
    ref = ops.gemm_bn_cat_f16(a, x, wcat, shift, relu=True)
    got = ops.gemm_bn_cat_f16(y, x, wcat, shift, relu=True, a1_ss=ss)
    assert torch.equal(got, ref)


def test_engine_without_a2_in_hbm_agrees_with_the_materialising_plan(bu):
    """MCEngine.fuse_a2 on / off on the multimodal net at 256 x 256: the only arithmetic difference is the summation order
    of conv3's first-moment column sums (fp32), i.e. BatchNorm statistics equal to ~1e-7 relative; logits must agree far
    inside the fp16-operand tolerance of DESIGN 4.3 (1.5e-3 of scale against the fp32 oracle)."""
    import bnn_oracle as O
    from mauv.engine import MCEngine
    _, model = bu.build_pair("multimodal")
    img, bathy, sss, _ = O.synthetic_batch(3, size=256)
    xs = [t.cuda() for t in (img, bathy, sss)]
    eng = MCEngine(model)
    assert eng.fuse_a2
    a = eng.forward_mc(xs, 4, seed=5, group=4, sample0=0)
    assert torch.equal(a, eng.forward_mc(xs, 4, seed=5, group=3, sample0=0))
    eng.fuse_a2 = False
    b = eng.forward_mc(xs, 4, seed=5, group=4, sample0=0)
    scale = b.abs().max().item()
    err = (a - b).abs().max().item() / scale
    print(f"fuse_a2 on/off: max |dlogit| / scale = {err:.2e}")
    assert err < 1.5e-3, err
    assert torch.equal(a.mean(0).argmax(-1), b.mean(0).argmax(-1))


@pytest.mark.parametrize("G,B,H,Cin,Cout,k,stride", [(2, 3, 12, 128, 128, 3, 1), (3, 4, 32, 128, 128, 3, 1), (2, 2, 32, 128, 128, 3, 2),
                                                      (2, 5, 16, 256, 128, 1, 2), (1, 3, 20, 128, 96, 3, 1)])
def test_m_stacked_tiles_of_the_n128_im2col_convs(bu, G, B, H, Cin, Cout, k, stride):
    """gemm_f16_tc_kernel<256, 0, 3> (two 128-row A tiles share one B tile; taken by im2col convs with 64 < N <= 128 and
    K >= 512) against F.conv2d on the same fp16 operands: outputs within the fp16 rounding of the result, per-channel
    statistics equal to the sums of the STORED values; ragged M (second half tile partial or empty), strides, N < 128."""
    import torch.nn.functional as F
    from mauv import ops
    torch.manual_seed(G * 7 + B)
    x = (torch.randn(G * B, H, H, Cin, device="cuda") * 0.5).half()
    w = (torch.randn(G, Cout, k, k, Cin, device="cuda") * 0.1).half()
    pad = k // 2
    y, st = ops.conv2d_im2col_f16(x, w.view(G, Cout, -1), G, k, k, stride, pad, stats=True)
    refs = []
    for g in range(G):
        xg = x[g * B:(g + 1) * B].float().permute(0, 3, 1, 2)
        refs.append(F.conv2d(xg, w[g].float().permute(0, 3, 1, 2), None, stride, pad).permute(0, 2, 3, 1))
    ref = torch.cat(refs)
    assert y.shape == ref.shape
    assert (y.float() - ref).abs().max().item() <= 1e-3 * ref.abs().max().item() + 1e-3
    yv = y.float().view(G, -1, Cout)
    M = yv.shape[1]
    assert st.shape == (G, (M + 127) // 128, Cout, 2)
    s = st.double().sum(1)
    assert torch.allclose(s[..., 0], yv.double().sum(1), rtol=1e-5, atol=1e-2)
    assert torch.allclose(s[..., 1], (yv.double() ** 2).sum(1), rtol=1e-5, atol=1e-2)
    # per-tile partials: tile t covers rows 128 t .. 128 t + 127
    t_last = (M + 127) // 128 - 1
    assert torch.allclose(st[:, t_last, :, 0].double(), yv[:, 128 * t_last:].double().sum(1), rtol=1e-5, atol=1e-2)


@pytest.mark.parametrize("B,C,size", [(2, 3, 256), (3, 1, 256), (2, 3, 64), (1, 1, 40)])
def test_row_tiled_stem_im2col_matches_unfold(bu, B, C, size):
    """stem_im2col_rows_kernel (7x7 / 2 / pad 3 stems: input rows staged in shared memory) against F.unfold in the (r, s, c)
    K order of the NHWC implicit GEMM; the K padding columns are zero."""
    import torch.nn.functional as F
    from mauv import ops
    torch.manual_seed(C * 10 + B)
    x = torch.randn(B, C, size, size, device="cuda")
    a = ops.stem_im2col_f16(x, 7, 7, 2, 3)
    cols = F.unfold(x, 7, padding=3, stride=2)                # [B, C*49, L] in (c, r, s) order
    L = cols.shape[2]
    ref = cols.view(B, C, 49, L).permute(0, 3, 2, 1).reshape(B * L, 49 * C).half()
    assert a.shape == (B * L, (49 * C + 7) // 8 * 8)
    assert torch.equal(a[:, :49 * C], ref)
    assert (a[:, 49 * C:] == 0).all()


@pytest.mark.parametrize("G,M,N,K,res,xf", [(3, 8192, 512, 128, True, False), (4, 8192, 256, 64, True, True), (3, 12800, 512, 128, True, True),
                                            (5, 4096 + 64, 1024, 128, False, False)])
def test_resident_weight_tile_tails_against_a_torch_reference(bu, G, M, N, K, res, xf):
    """Fused-BN tails large enough (>= 148 tiles) to take the resident-weight layout (one B load per sample and CTA, the tile is
    replaced when the CTA's tile sequence crosses into the next sample): out = relu((A W^T) * scale + shift + residual) against
    fp32 torch on the same fp16 operands, with and without the operand transform; ragged last m-tile; N = 1024 (4 n-tiles)."""
    from mauv import ops
    y, ss = _raw_and_bn(G, M, K, 17 + N)
    w = (torch.randn(G, N, K, device="cuda") * 0.05).half()
    ss3 = torch.stack([torch.rand(G, N, device="cuda") + 0.5, torch.randn(G, N, device="cuda") * 0.2], dim=-1).contiguous()
    r = torch.randn(G, M, N, device="cuda").half() if res else None
    if xf:
        a = ops.bn_act_f16(y, ss, G, K, relu=True)
        got = ops.gemm_bn_act_f16(y, w, ss3, residual=r, relu=True, a_ss=ss)
    else:
        a = y
        got = ops.gemm_bn_act_f16(a, w, ss3, residual=r, relu=True)
    ref = torch.bmm(a.float(), w.float().transpose(1, 2)) * ss3[:, None, :, 0] + ss3[:, None, :, 1]
    if res:
        ref = ref + r.float()
    ref = torch.relu(ref)
    err = (got.float() - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-3, err
    # every sample used ITS weights: permuting the samples permutes the outputs
    perm = torch.arange(G - 1, -1, -1, device="cuda")
    got_p = ops.gemm_bn_act_f16(a[perm].contiguous() if not xf else y[perm].contiguous(), w[perm].contiguous(), ss3[perm].contiguous(),
                                residual=None if r is None else r[perm].contiguous(), relu=True,
                                a_ss=ss[perm].contiguous() if xf else None)
    assert torch.equal(got_p, got[perm])


@pytest.mark.parametrize("G,B,fin,fout,bias", [(20, 130, 2048, 128, True), (3, 256, 384, 1263, True), (3, 128, 96, 7, False)])
def test_head_linear_128_row_tiles_equal_the_32_row_tiles(bu, G, B, fin, fout, bias):
    """sampled_linear_kernel<128> (batches >= 128 rows: the sampled weight tile is reused by 128 rows) against the 32-row
    instance run on row slices < 128 - same Philox ids, same accumulation order: torch.equal - and against fp64."""
    from mauv import ops
    torch.manual_seed(fin + B)      # (the 128-row instance is taken when its grid still fills the GPU: cases 1 and 2)
    x = torch.randn(G, B, fin, device="cuda")
    mu, rho = torch.randn(fout, fin, device="cuda") * 0.05, torch.full((fout, fin), -3.0, device="cuda") + torch.randn(fout, fin, device="cuda") * 0.1
    mu_b = torch.randn(fout, device="cuda") * 0.1 if bias else None
    rho_b = torch.full((fout,), -3.0, device="cuda") if bias else None
    kw = dict(seed=99, layer_id=5, sample0=11)
    big = ops.sampled_linear_f32(x, mu, rho, mu_b, rho_b, **kw)
    parts = [ops.sampled_linear_f32(x[:, lo:lo + 100].contiguous(), mu, rho, mu_b, rho_b, **kw) for lo in range(0, B, 100)]
    assert torch.equal(big, torch.cat(parts, dim=1))
    eps = torch.randn(G, fout, fin, device="cuda")
    got = ops.sampled_linear_f32(x, mu, rho, None, None, eps_w=eps)
    w = mu.double() + torch.log1p(torch.exp(rho.double())) * eps.double()
    ref = torch.einsum("gbi,goi->gbo", x.double(), w)
    assert (got.double() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item() + 1e-5


def test_three_trunks_on_three_streams_give_bit_identical_logits(bu):
    """MCEngine with the trunks forked onto three streams (the plan of the small sample groups at 4 / 8 GPUs) against the
    sequential walk, eagerly and inside a CUDA-graph replay of the predictor: torch.equal."""
    import bnn_oracle as O
    from mauv.engine import MCEngine
    from mauv.inference.predictors import MCPredictor
    _, model = bu.build_pair("multimodal")
    img, bathy, sss, _ = O.synthetic_batch(3, size=256)
    xs = [t.cuda() for t in (img, bathy, sss)]
    eng = MCEngine(model)
    eng.trunk_streams = False
    seq = eng.forward_mc(xs, 4, seed=21, group=4, sample0=0)
    eng.trunk_streams = True
    par = eng.forward_mc(xs, 4, seed=21, group=4, sample0=0)
    torch.cuda.synchronize()
    assert torch.equal(par, seq)
    assert torch.equal(eng.forward_mc(xs, 4, seed=21, group=2, sample0=0), seq)
    pred = MCPredictor(model, 4, use_graph=True)
    pred.engine.trunk_streams = True
    from mauv.bayesian import manual_seed
    manual_seed(21)
    a = pred.mc_logits(xs, sample0=0)            # records the graph (after an eager warm-up), then replays
    b = pred.mc_logits(xs, sample0=0)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(a, seq)


def test_train_step_with_trunks_on_three_streams_is_bit_identical(bu):
    """TrainEngine.step of the multimodal net with the three trunks (forward AND backward) on three streams against the
    sequential walk: loss, every gradient and the BN running statistics torch.equal - eagerly and as a recorded CUDA graph."""
    import bnn_oracle as O
    from mauv.bayesian import manual_seed
    from mauv.train_engine import TrainEngine
    _, model = bu.build_pair("multimodal")
    state0 = {k: v.clone() for k, v in model.state_dict().items()}
    batches = [O.synthetic_batch(2, seed=70 + i, size=64) for i in range(3)]
    manual_seed(13)

    def run(eng, i, sample0):
        model.load_state_dict(state0)
        for p in model.parameters():
            if p.grad is not None:
                p.grad.zero_()
        img, bathy, sss, labels = batches[i]
        res = eng.step([img.cuda(), bathy.cuda(), sss.cuda()], labels, 3, 1e-4, sample0=sample0)
        torch.cuda.synchronize()
        return ({n: p.grad.clone() for n, p in model.named_parameters()}, res["loss"].item(),
                {k: v.clone() for k, v in model.state_dict().items() if "running" in k})

    seq = TrainEngine(model)
    seq.use_graph = False
    seq.trunk_streams = False
    ref = run(seq, 2, 7)
    par = TrainEngine(model)
    par.use_graph = False
    par.trunk_streams = True
    got = run(par, 2, 7)
    assert got[1] == ref[1]
    for n in ref[0]:
        assert torch.equal(got[0][n], ref[0][n]), n
    for k in ref[2]:
        assert torch.equal(got[2][k], ref[2][k]), k
    gr = TrainEngine(model)              # default: trunk streams on, graph replay from the third call
    assert gr.use_graph and gr._train_trunks_parallel()
    run(gr, 0, 3)
    run(gr, 1, 5)
    got = run(gr, 2, 7)
    assert got[1] == ref[1]
    for n in ref[0]:
        assert torch.equal(got[0][n], ref[0][n]), n
