"""Host-side logic that needs no GPU: MC-sample sharding and the world_size-2 gather (gloo, CPU tensors)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def test_shard_samples_partitions_exactly():
    from mauv.inference.predictors import shard_samples
    for S in (1, 5, 30, 31, 100):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_samples(S, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == S
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == S
    assert [shard_samples(30, 8, r) for r in range(8)][:2] == [(0, 4), (4, 8)]
    assert shard_samples(30, 4, 3) == (23, 30)      # 8+8+7+7


def _worker(rank, world, port, S, B, C, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    from mauv.inference.predictors import gather_sample_blocks, shard_samples
    full = torch.arange(S * B * C, dtype=torch.float32).view(S, B, C)
    lo, hi = shard_samples(S, world, rank)
    got = gather_sample_blocks(full[lo:hi].clone() if hi > lo else None, S, B, C, world, "cpu")
    q.put((rank, bool(torch.equal(got, full))))
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("S", [5, 1])
def test_gather_sample_blocks_world2_gloo(S):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, S, 3, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


# ------------------------------------------------------------------ training: flat gradient buffer + data-parallel mean
def _small_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))


def test_flat_grads_views_zero_and_reattach():
    from mauv.flatgrad import FlatGrads
    m = _small_model()
    m(torch.randn(4, 5)).sum().backward()
    before = [p.grad.clone() for p in m.parameters()]
    fg = FlatGrads(m.parameters())
    assert all(torch.equal(p.grad, b) for p, b in zip(m.parameters(), before))          # existing gradients are kept
    assert all(p.grad.data_ptr() >= fg.flat.data_ptr() for p in m.parameters())
    assert all((p.grad.data_ptr() - fg.flat.data_ptr()) % 128 == 0 for p in m.parameters())
    m(torch.randn(4, 5)).sum().backward()                                               # autograd accumulates into the views
    assert fg.flat.abs().sum() > 0 and bool(fg.finite())
    torch.optim.SGD(m.parameters(), lr=0.1).zero_grad(set_to_none=True)
    assert all(p.grad is None for p in m.parameters())
    fg.ensure_attached()
    assert all(p.grad is not None and float(p.grad.abs().sum()) == 0.0 for p in m.parameters())
    next(m.parameters()).grad.fill_(float("nan"))
    assert not bool(fg.finite())
    fg.zero()
    assert float(fg.flat.abs().sum()) == 0.0


def test_group_splits_fill_the_gpu():
    from mauv.train_engine import _group_splits
    assert _group_splits(32768, 30, 256, 64) == 8            # layer1 conv3 at cfg3: 30 samples x 2 m-tiles x 1 n-tile = 60 tiles
    assert _group_splits(32768, 30, 64, 576) == 4            # layer1 conv2
    assert _group_splits(512, 30, 512, 4608) == 1            # layer4 conv2: already 2 160 tiles
    assert _group_splits(8, 2, 64, 64) == 1                  # tiny test shapes never split below 2 048 pixels


def test_train_engine_selection_rules():
    """train_*_model take the S-batched engine only for plain mean CrossEntropyLoss; on a machine without the device the
    selection returns None (the layer path then refuses CPU tensors loudly - there is no CPU fallback)."""
    from mauv.train.multimodal import train_engine_for
    m = _small_model()
    assert train_engine_for(m, torch.nn.MSELoss()) is None
    assert train_engine_for(m, torch.nn.CrossEntropyLoss(label_smoothing=0.1)) is None
    assert train_engine_for(m, torch.nn.CrossEntropyLoss(reduction="sum")) is None
    if not torch.cuda.is_available():
        assert train_engine_for(m, torch.nn.CrossEntropyLoss()) is None


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    from mauv.flatgrad import FlatGrads
    m = _small_model()
    fg = FlatGrads(m.parameters())
    x = torch.arange(8 * 5, dtype=torch.float32).view(8, 5) / 10
    y = torch.arange(8) % 3
    shard = slice(rank * 4, rank * 4 + 4)
    torch.nn.functional.cross_entropy(m(x[shard]), y[shard]).backward()                 # local shard gradient into the views
    fg.all_reduce_mean()
    ref = _small_model()
    torch.nn.functional.cross_entropy(ref(x), y).backward()                              # global-minibatch gradient
    ok = all(torch.allclose(p.grad, r.grad, atol=1e-6) for p, r in zip(m.parameters(), ref.parameters()))
    q.put((rank, ok))
    torch.distributed.destroy_process_group()


def test_data_parallel_gradient_mean_world2_gloo():
    """The N > 1 training exchange: shard gradients averaged by one all-reduce over the flat buffer == the gradient of the
    global minibatch (equal shards)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_load_reference_weights_key_normalisation(tmp_path):
    """Checkpoints as the reference writes / publishes them (DataParallel `module.` prefix, trunks one `.model.` level deeper,
    7-class output layer) load into the mauv multimodal model: Examples/Example_Inference_model.py:82-112."""
    import bnn_oracle as O
    from mauv.models.model_utils import define_models, load_reference_weights
    torch.manual_seed(3)
    src = define_models(torch.device("cpu"), 7, O.DEFAULT_PRIOR)["multimodal_model"]
    published = {}
    for k, v in src.state_dict().items():
        for b in ("image_model_feat", "bathy_model_feat", "sss_model_feat"):
            if k.startswith(b + "."):
                k = b + ".model." + k[len(b) + 1:]
                break
        published["module." + k] = v.clone()
    path = tmp_path / "pytorch_model.bin"
    torch.save(published, path)
    torch.manual_seed(4)
    dst = define_models(torch.device("cpu"), 7, O.DEFAULT_PRIOR)["multimodal_model"]
    missing, unexpected = load_reference_weights(dst, str(path))
    assert missing == [] and unexpected == []
    assert all(torch.equal(v, dst.state_dict()[k]) for k, v in src.state_dict().items())
    # 5-class head: fc2 keeps its fresh initialisation, everything else is loaded
    dst5 = define_models(torch.device("cpu"), 5, O.DEFAULT_PRIOR)["multimodal_model"]
    fc2_before = dst5.fc2.mu_weight.detach().clone()
    missing, unexpected = load_reference_weights(dst5, published, num_classes=5)
    assert sorted(missing) == sorted(k for k in dst5.state_dict() if k.startswith("fc2.")) and unexpected == []
    assert torch.equal(dst5.fc2.mu_weight, fc2_before)
    assert torch.equal(dst5.fc1.mu_weight, src.fc1.mu_weight)


def test_bench_roofline_launch_model():
    """bench.py's per-launch roofline model: algorithmic flops and minimum HBM bytes parsed from the profiling tags."""
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("bench_mod", Path(__file__).resolve().parent.parent / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    f, b = bench.tc_launch_model("mauv_gemm_f16", "G10 M1048576 N256 K64")
    assert f == 2.0 * 10 * 1048576 * 256 * 64
    assert b == 10 * 1048576 * 64 * 2 + 10 * 256 * 64 * 2 + 10 * 1048576 * 256 * 2          # A + W + Y
    f, b = bench.tc_launch_model("mauv_conv2d_im2col_f16", "G10 M262144 N128 K1152 3x3/2")
    assert f == 2.0 * 10 * 262144 * 128 * 1152
    assert b == 10 * 262144 * 4 * 128 * 2 + 10 * 128 * 1152 * 2 + 10 * 262144 * 128 * 2       # stride 2: 4x the input pixels
    f, b = bench.tc_launch_model("mauv_gemm_bn_f16", "stats G10 M65536 N1024 K256")
    assert b == 10 * 65536 * 256 * 2 + 10 * 1024 * 256 * 2                                     # statistics pass: no output
    f, b = bench.tc_launch_model("mauv_gemm_bn_f16", "fused G10 M65536 N1024 K256 res1")
    assert b == 10 * 65536 * 256 * 2 + 10 * 1024 * 256 * 2 + 2 * 10 * 65536 * 1024 * 2         # + residual read + output
    f, b = bench.tc_launch_model("mauv_gemm_f16", "G10 M4194304 N64 K152")
    assert b == 4194304 * 152 * 2 + 10 * 64 * 152 * 2 + 10 * 4194304 * 64 * 2                  # stem: A shared by all samples
    f, b = bench.tc_launch_model("mauv_conv3x3_c64_f16", "G10 M1048576 N64 K576 stream")
    assert b == 10 * 1048576 * 64 * 2 * 2 + 10 * 64 * 576 * 2
    assert bench.tc_launch_model("mauv_bn_act_f16", "G10 M100 C64 res0 dual0") is None
    # round 3: the stem kernel (A shared, only the pooled quarter is written), the operand-transform tails, the second moments
    f, b = bench.tc_launch_model("mauv_stem_conv_pool_f16", "G15 M4194304 N64 K152 pool")
    assert f == 2.0 * 15 * 4194304 * 64 * 152
    assert b == 4194304 * 152 * 2 + 15 * 64 * 152 * 2 + 15 * (4194304 // 4) * 64 * 2
    f, b = bench.tc_launch_model("mauv_gemm_bn_xf_f16", "fused G30 M1048576 N256 K64 res1 xf")
    assert b == 30 * 1048576 * 64 * 2 + 30 * 256 * 64 * 2 + 2 * 30 * 1048576 * 256 * 2
    f, b = bench.tc_launch_model("mauv_gram_bn_f16", "G30x256 Cout64 K64 px4096 xf")
    assert f == 2.0 * 30 * 256 * 64 * 64 * 4096 and b == 30 * 256 * 4096 * 64 * 2 + 30 * 256 * 64 * 64 * 4
    f, b = bench.tc_launch_model("mauv_gemm_bn_cat_xf_f16", "fused G30 M262144 N512 K384 cat xf")
    assert f == 2.0 * 30 * 262144 * 512 * 384


def test_every_tcgen05_entry_point_is_bound_and_counted_by_the_bench():
    """bench.py's roofline family is ops.TCGEN05_ENTRY_POINTS: every name must be a bound C-ABI symbol (a kernel left out of
    the family would inflate the reported fraction: its FLOPs counted, its time not)."""
    from pathlib import Path
    from mauv import _lib, ops
    for name in ops.TCGEN05_ENTRY_POINTS:
        assert name in _lib.SIGNATURES, name
    src = (Path(__file__).resolve().parent.parent / "multimodal-auv_b200" / "mauv" / "ops.py").read_text()
    hot = {"mauv_gemm_f16", "mauv_conv2d_im2col_f16", "mauv_gemm_bn_f16", "mauv_gemm_bn_xf_f16", "mauv_gemm_bn_cat_f16",
           "mauv_gemm_bn_cat_xf_f16", "mauv_conv3x3_c64_f16", "mauv_wgrad_f16", "mauv_gram_bn_f16", "mauv_stem_conv_pool_f16"}
    assert hot <= set(ops.TCGEN05_ENTRY_POINTS)
    for name in hot:
        assert f'"{name}"' in src


# ------------------------------------------------------------------ a10: the product's own dnn_to_bnn + MOPED vs the oracle
def test_product_dnn_to_bnn_moped_equals_oracle_and_reference_checksum():
    """mauv.models.model_utils.define_models -> mauv.bayesian.dnn_to_bnn (the path bench.py and users take) against the
    oracle's restatement of bayesian-torch's converter on the same seeded torchvision nets: every parameter and buffer
    bit-equal, and the parameter checksum equal to the one recorded from the REFERENCE's define_models (tests/golden)."""
    import logging
    import bnn_oracle as O
    from mauv.bayesian import Conv2dReparameterization, LinearReparameterization
    from mauv.models.model_utils import define_models
    logging.disable(logging.WARNING)
    try:
        torch.manual_seed(1234)
        prod = define_models(torch.device("cpu"), 7, dict(O.DEFAULT_PRIOR))
    finally:
        logging.disable(logging.NOTSET)
    orac = O.define_models(7, seed=1234, unimodal=True)
    for key in ("multimodal_model", "image_model", "bathy_model", "sss_model"):
        ps, os_ = prod[key].state_dict(), orac[key].state_dict()
        assert list(ps.keys()) == list(os_.keys()), key
        for k in ps:
            assert torch.equal(ps[k], os_[k]), (key, k)
    mm = prod["multimodal_model"]
    layers = [m for m in mm.modules() if isinstance(m, (Conv2dReparameterization, LinearReparameterization))]
    assert len(layers) == 174 and all(l.dnn_to_bnn_flag for l in layers)
    assert not any(isinstance(m, (torch.nn.Conv2d, torch.nn.Linear)) for m in mm.modules())
    # MOPED: mu = w, softplus(rho) = delta * |w| (delta = 0.1); rho is NOT the constant posterior_rho_init
    conv = mm.image_model_feat.layer2[0].conv2
    sigma = torch.log1p(torch.exp(conv.rho_kernel.double()))
    assert torch.allclose(sigma, 0.1 * conv.mu_kernel.double().abs(), rtol=1e-4, atol=1e-12)
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_small.pt"), weights_only=False)
    assert abs(float(sum(p.detach().double().sum() for p in mm.parameters())) - gold["param_checksum"]) < 1e-6
    # without MOPED the posterior starts at N(posterior_mu_init, 0.1) / N(posterior_rho_init, 0.1) like bayesian-torch's layers
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.Linear(8, 4))
    from mauv.bayesian import dnn_to_bnn
    dnn_to_bnn(net, dict(O.DEFAULT_PRIOR, moped_enable=False))
    assert abs(float(net[0].rho_kernel.mean()) + 3.0) < 0.1 and abs(float(net[1].mu_weight.mean())) < 0.1


def test_load_models_and_pathlike_checkpoints(tmp_path):
    """models/model_utils.py:66-101 mirror: existing paths are loaded, missing ones only warn; load_reference_weights takes
    a pathlib.Path (what the reference passes around)."""
    import logging
    import pathlib
    from mauv.models.model_utils import load_models, load_pretrained_resnet_as_feature_extractor, load_reference_weights
    logging.disable(logging.ERROR)
    try:
        torch.manual_seed(3)
        src = load_pretrained_resnet_as_feature_extractor(input_channels=1)
        p = tmp_path / "sss.pth"
        torch.save(src.state_dict(), p)
        img, chan, sss = load_models({"image": str(tmp_path / "missing.pth"), "sss": str(p)}, torch.device("cpu"), 7)
        assert all(torch.equal(a, b) for a, b in zip(sss.state_dict().values(), src.state_dict().values()))
        assert img.conv1.in_channels == 3 and chan.conv1.in_channels == 3 and sss.conv1.in_channels == 1
        lin = torch.nn.Linear(2, 2)
        ck = tmp_path / "lin.pth"
        torch.save({"module." + k: v for k, v in lin.state_dict().items()}, ck)
        missing, unexpected = load_reference_weights(torch.nn.Linear(2, 2), pathlib.Path(ck))
        assert missing == [] and unexpected == []
        with pytest.raises(TypeError):
            load_reference_weights(lin, 12345)
    finally:
        logging.disable(logging.NOTSET)


class _FakeEngine:
    """Stands in for TrainEngine in the CPU test of the drivers' gradient exchange: same flatten / all-reduce plumbing
    (mauv.flatgrad.FlatGrads), gradients written straight into `.grad` like the engine does."""

    def __init__(self, module):
        self.module, self._flat = module, None

    def flatten_grads(self):
        from mauv.flatgrad import FlatGrads
        self._flat = FlatGrads(self.module.parameters())

    def allreduce_grads(self, group=None):
        self._flat.all_reduce_mean(group)


def _ddp_sync_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    from mauv.train.multimodal import ddp_sync_group
    m = _small_model()
    ddp = torch.nn.parallel.DistributedDataParallel(m)
    eng = _FakeEngine(m)
    assert ddp_sync_group(m, eng) is False                      # not wrapped: nothing to exchange
    assert ddp_sync_group(ddp, None) is False                   # layer path: DDP's own reducer does it
    group = ddp_sync_group(ddp, eng)
    assert group is not False and eng._flat is not None
    for p in m.parameters():                                    # the "engine" writes rank-dependent gradients into .grad
        p.grad.fill_(float(rank + 1))
    eng.allreduce_grads(group)
    ok = all(torch.allclose(p.grad, torch.full_like(p.grad, 1.5)) for p in m.parameters())
    q.put((rank, ok))
    torch.distributed.destroy_process_group()


def test_drop_in_train_loops_average_engine_gradients_under_ddp_world2_gloo():
    """TrainEngine bypasses DistributedDataParallel's reducer, so train_*_model must all-reduce the flat
    gradient buffer themselves when handed a DDP-wrapped model (mauv.train.multimodal.ddp_sync_group)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_batch_slice_partitions_rows_for_the_sliced_upload():
    """8f-2 staging: every rank uploads rows [lo, hi) and the slices tile the batch exactly; chunk is the padded length the
    all-gather uses (equal on all ranks)."""
    from mauv.inference.predictors import batch_slice
    for B in (1, 7, 8, 255, 256, 257):
        for world in (1, 2, 3, 4, 8):
            rows, chunks = [], set()
            for r in range(world):
                lo, hi, chunk = batch_slice(B, world, r)
                assert 0 <= lo <= hi <= B and hi - lo <= chunk
                rows += list(range(lo, hi))
                chunks.add(chunk)
                assert lo == min(r * chunk, B)          # gathered position of the slice = r * chunk
            assert rows == list(range(B)) and len(chunks) == 1 and chunks.pop() * world >= B
