"""Kernel micro-benchmarks on the B200 box (CUDA events, L2 flushed by the working-set size).
  python tests/gpu_microbench.py gemm  G M N K [iters]
  python tests/gpu_microbench.py conv  G B H W Cin Cout k stride pad [iters]
  python tests/gpu_microbench.py bnact G M C [iters]
  python tests/gpu_microbench.py mcreduce S B C [iters]
  python tests/gpu_microbench.py wgrad G B H W Cin Cout k stride pad [splits]
  python tests/gpu_microbench.py bnbwd G M C
  python tests/gpu_microbench.py kl | sample | hbm | hbmwrite
  python tests/gpu_microbench.py layers        (the ResNet-50 trunk shapes at B=256)
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "multimodal-auv_b200"))
import torch

from mauv import ops

dev = "cuda"


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def gemm(G, M, N, K, iters=5, shared=False, quiet=False):
    a = torch.randn((M, K) if shared else (G, M, K), device=dev, dtype=torch.float16)
    w = torch.randn(G, N, K, device=dev, dtype=torch.float16)
    y = torch.empty(G, M, N, device=dev, dtype=torch.float16)
    st = torch.empty(G, ops.gemm_m_tiles(M), N, 2, device=dev)
    ms = timeit(lambda: ops.gemm_f16(a, w, stats=True, shared_a=shared, out=y, stats_out=st), iters)
    fl = 2.0 * G * M * N * K
    by = (a.numel() + y.numel() + w.numel()) * 2 + st.numel() * 4
    if not quiet:
        print(f"gemm G={G} M={M} N={N} K={K} shared={shared}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e9:.2f} TB/s", flush=True)
    return ms


def gemm_bn(G, M, N, K, iters=5):
    """conv3 recompute scheme: statistics pass + fused BN/residual/ReLU pass."""
    a = torch.randn(G, M, K, device=dev, dtype=torch.float16)
    w = torch.randn(G, N, K, device=dev, dtype=torch.float16)
    r = torch.randn(G, M, N, device=dev, dtype=torch.float16)
    ss = torch.rand(G, N, 2, device=dev)
    out = torch.empty(G, M, N, device=dev, dtype=torch.float16)
    st = torch.empty(G, (M + 255) // 256, N, 2, device=dev)
    ms = timeit(lambda: ops.gemm_stats_f16(a, w, stats_out=st), iters)
    print(f"gemm_bn stats pass G={G} M={M} N={N} K={K}: {ms:.3f} ms  {a.numel() * 2 / ms / 1e9:.2f} TB/s (reads A only)", flush=True)
    ms = timeit(lambda: ops.gemm_bn_act_f16(a, w, ss, residual=r, relu=True, out=out), iters)
    by = (a.numel() + 2 * out.numel()) * 2
    print(f"gemm_bn fused pass G={G} M={M} N={N} K={K}: {ms:.3f} ms  {by / ms / 1e9:.2f} TB/s (A + residual + out)", flush=True)


def conv(G, B, H, W, Cin, Cout, k, stride, pad, iters=5):
    x = torch.randn(G * B, H, W, Cin, device=dev, dtype=torch.float16)
    w = torch.randn(G, Cout, k * k * Cin, device=dev, dtype=torch.float16)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    y = torch.empty(G * B, Ho, Wo, Cout, device=dev, dtype=torch.float16)
    st = torch.empty(G, ops.gemm_m_tiles(B * Ho * Wo), Cout, 2, device=dev)
    ms = timeit(lambda: ops.conv2d_im2col_f16(x, w, G, k, k, stride, pad, stats=True, out=y, stats_out=st), iters)
    if Cin == 64 and Cout == 64 and k == 3 and stride == 1:
        ms2 = timeit(lambda: ops.conv3x3_c64_f16(x, w, G, stats=True), iters)
        print(f"   padded-stream kernel: {ms2:.3f} ms ({ms / ms2:.2f}x the im2col path)", flush=True)
    fl = 2.0 * G * B * Ho * Wo * Cout * k * k * Cin
    by = (x.numel() + y.numel() + w.numel()) * 2 + st.numel() * 4
    print(f"conv G={G} B={B} {H}x{W} {Cin}->{Cout} k{k}/{stride}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e9:.2f} TB/s", flush=True)
    return ms


def bnact(G, M, C, iters=5):
    y = torch.randn(G, M, C, device=dev, dtype=torch.float16)
    r = torch.randn(G, M, C, device=dev, dtype=torch.float16)
    ss = torch.rand(G, C, 2, device=dev)
    out = torch.empty_like(y)
    for name, kw, nbuf in (("plain", {}, 2), ("residual", {"residual": r}, 3), ("dual", {"y2": r, "ss2": ss}, 3)):
        ms = timeit(lambda: ops.bn_act_f16(y, ss, G, C, out=out, **kw), iters)
        print(f"bn_act {name} G={G} M={M} C={C}: {ms:.3f} ms  {y.numel() * 2 * nbuf / ms / 1e9:.2f} TB/s", flush=True)


def mcreduce(S, B, C, iters=5):
    lg = torch.randn(S, B, C, device=dev)
    ms = timeit(lambda: ops.mc_reduce(lg, 1e-8), iters)
    by = S * B * C * 4 + B * (2 * C + 4) * 4 + B * 16
    print(f"mc_reduce S={S} B={B} C={C}: {ms:.4f} ms  {by / ms / 1e6:.1f} GB/s (algorithmic bytes {by / 1e6:.1f} MB)", flush=True)


def kl():
    shapes = [(2048, 512, 1, 1)] * 40 + [(512, 512, 3, 3)] * 20
    mus = [torch.randn(s, device=dev) * 0.05 for s in shapes]
    rhos = [torch.randn(s, device=dev) - 4 for s in shapes]
    gr = [(torch.zeros_like(m), torch.zeros_like(m)) for m in mus]
    n = sum(m.numel() for m in mus)
    plan = ops.KlPlan(list(zip(mus, rhos)), dev)
    ms = timeit(lambda: plan.run(0.0, 1.0), 5)
    print(f"kl fwd n={n}: {ms:.3f} ms  {n * 8 / ms / 1e6:.1f} GB/s (8 B/param)", flush=True)
    plan2 = ops.KlPlan(list(zip(mus, rhos)), dev, grads=gr)
    ms = timeit(lambda: plan2.run(0.0, 1.0, grad_scale=1e-3), 5)
    print(f"kl fwd+bwd n={n}: {ms:.3f} ms  {n * 24 / ms / 1e6:.1f} GB/s (24 B/param)", flush=True)


def sample():
    for shape, G in (((512, 512, 3, 3), 10), ((2048, 1024, 1, 1), 10), ((64, 64, 3, 3), 10)):
        mu = torch.randn(shape, device=dev) * 0.05
        rho = torch.randn(shape, device=dev) - 4
        n = mu.numel()
        out = torch.empty(G, shape[0], n // shape[0], device=dev, dtype=torch.float16)
        ms = timeit(lambda: ops.sample_weights_f16(mu, rho, G, seed=1, layer_id=3, out=out), 5)
        print(f"sample_weights {shape} G={G} philox: {ms:.3f} ms  {(n * 8 + G * n * 2) / ms / 1e6:.1f} GB/s "
              f"({G * n / ms / 1e6:.1f} Gweights/s)", flush=True)


def wgrad(G, B, H, W, Cin, Cout, k, stride, pad, splits=1, iters=5):
    """weight gradient straight from the NHWC tensors (MN-major tcgen05 operands)"""
    x = torch.randn(G * B, H, W, Cin, device=dev, dtype=torch.float16)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    dy = torch.randn(G * B, Ho, Wo, Cout, device=dev, dtype=torch.float16)
    ms = timeit(lambda: ops.wgrad_f16(dy, x, G, splits, k, k, stride, pad), iters)
    fl = 2.0 * G * B * Ho * Wo * Cout * k * k * Cin
    by = (x.numel() + dy.numel() + G * splits * Cout * k * k * Cin) * 2
    print(f"wgrad G={G} B={B} {H}x{W} {Cin}->{Cout} k{k}/{stride} splits={splits}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  "
          f"{by / ms / 1e9:.2f} TB/s", flush=True)


def gram(G, M, K, N=0, iters=5):
    """closed-form BN statistics: second-moment contraction a^T a (tcgen05 weight-gradient kernel, gram mode) + evaluation"""
    N = N or 4 * K
    a = torch.relu(torch.randn(G, M, K, device=dev, dtype=torch.float16))
    w = (torch.randn(G, N, K, device=dev) * 0.05).half()
    cs = ops.colsum_f16(a, G, K)
    gamma, beta = torch.ones(N, device=dev), torch.zeros(N, device=dev)
    sp = ops.gram_splits(M, G, K)
    x4 = a.view(G, M, 1, K)
    ms = timeit(lambda: ops.wgrad_f16(x4, x4, G, sp, 1, 1, 1, 0), iters)
    print(f"gram a^T a G={G} M={M} K={K} splits={sp}: {ms:.3f} ms  {a.numel() * 2 / ms / 1e9:.2f} TB/s (reads a once)  "
          f"{2.0 * G * M * K * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    ms2 = timeit(lambda: ops.bn_stats_from_gram(a, cs, w, M, gamma, beta, 1e-5, 0.0), iters)
    print(f"   contraction + evaluation (N={N}): {ms2:.3f} ms; statistics pass it replaces: "
          f"{timeit(lambda: ops.gemm_stats_f16(a, w), iters):.3f} ms", flush=True)


def adam(n=146_608_278, iters=5):
    """fused Adam + finite guard over the flat buffers (n = the multimodal model's 2 x 73.3 M parameters)"""
    st = torch.zeros(8, dtype=torch.int32, device=dev)
    p, g, m, v = (torch.randn(n, device=dev) * s_ for s_ in (0.05, 1e-3, 0.0, 0.0))
    ms = timeit(lambda: ops.adam_step_f32(p, g, m, v, 1e-4, 0.9, 0.999, 1e-8, 0.0, st), iters)
    print(f"fused adam n={n}: {ms:.3f} ms  {n * 32 / ms / 1e6:.1f} GB/s (32 B / parameter: check 4 + update 28)", flush=True)
    params = [torch.nn.Parameter(torch.randn(k, device=dev)) for k in [n // 696] * 696]
    for q in params:
        q.grad = torch.randn_like(q) * 1e-3
    opt = torch.optim.Adam(params, lr=1e-4)
    ms2 = timeit(lambda: (all(bool(torch.isfinite(q.grad).all()) for q in params[:0]) or True) and opt.step(), iters)
    fin = timeit(lambda: torch.stack([torch.isfinite(q.grad).all() for q in params]).all().item(), 3)
    print(f"torch.optim.Adam over 696 tensors: {ms2:.3f} ms; per-parameter finite guard: {fin:.3f} ms", flush=True)


def bnbwd(G, M, C, iters=5):
    """one BatchNorm-backward site of a bottleneck tail (two upstream tensors, ReLU mask, dz output)"""
    y = torch.randn(G, M, C, device=dev, dtype=torch.float16)
    out = torch.relu(torch.randn(G, M, C, device=dev, dtype=torch.float16))
    d1 = torch.randn(G, M, C, device=dev, dtype=torch.float16)
    d2 = torch.randn(G, M, C, device=dev, dtype=torch.float16)
    bs = torch.stack([torch.zeros(G, C, device=dev), torch.ones(G, C, device=dev)], -1).contiguous()
    gamma = torch.ones(C, device=dev)
    gg, gb = torch.zeros(C, device=dev), torch.zeros(C, device=dev)

    def run():
        gs = ops.GradScratch(dev, 16)
        gs.f[0], gs.f[1] = 1.0, 2.0
        gs.used = 2
        ops.bn_bwd_site(gs, d1, gs.f.data_ptr(), y, bs, gamma, 1e-5, gg, gb, G, C, d2=d2, s2=gs.f.data_ptr() + 4,
                        relu_out=out, want_dz=True)
    ms = timeit(run, iters)
    n = y.numel() * 2
    print(f"bn_bwd site G={G} M={M} C={C} (reduce + coeffs + apply): {ms:.3f} ms  {(4 * n + 6 * n) / ms / 1e9:.2f} TB/s "
          f"(reduce reads 4 tensors; apply reads 4, writes 2)", flush=True)


def hbm():
    """HBM read-only / write-only / copy bandwidth with plain torch ops (measurement aid for the roofline split)."""
    n = 1 << 31
    a = torch.empty(n, dtype=torch.uint8, device=dev)
    b = torch.empty(n, dtype=torch.uint8, device=dev)
    ms = timeit(lambda: a.fill_(1), 5)
    print(f"write-only (fill 2 GiB): {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s", flush=True)
    af = a.view(torch.float32)
    ms = timeit(lambda: af.sum(), 5)
    print(f"read-only  (sum 2 GiB): {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s", flush=True)
    ms = timeit(lambda: b.copy_(a), 5)
    print(f"copy (2 GiB -> 2 GiB): {ms:.3f} ms  {2 * n / ms / 1e6:.1f} GB/s", flush=True)


def hbmwrite():
    """VERDICT r1 task 2(iii): is ~3.9 TB/s the write-only ceiling of HBM on this GPU? 4 GiB written five ways."""
    import ctypes
    from mauv import _lib
    lib = _lib.require_device()
    n = 1 << 32
    a = torch.empty(n, dtype=torch.uint8, device=dev)
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMemsetAsync.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    for name, fn in (("cudaMemsetAsync", lambda: rt.cudaMemsetAsync(a.data_ptr(), 0, n, st)),
                     ("torch fill_", lambda: a.fill_(1)),
                     ("st.global.v4", lambda: _lib.check(lib.mauv_membench_fill(a.data_ptr(), n, 7, 0, st))),
                     ("st.global.cs.v4", lambda: _lib.check(lib.mauv_membench_fill(a.data_ptr(), n, 7, 1, st))),
                     ("cp.async.bulk shared->global (TMA store engine)", lambda: _lib.check(lib.mauv_membench_fill(a.data_ptr(), n, 7, 2, st)))):
        ms = timeit(fn, 10, 3)
        rows.append((name, ms, n / ms / 1e6))
        print(f"write-only {name}: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s", flush=True)
    b = torch.empty(n // 2, dtype=torch.uint8, device=dev)
    ms = timeit(lambda: b.copy_(a[: n // 2]), 10, 3)
    print(f"copy 2 GiB -> 2 GiB: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s (read + write)", flush=True)
    ms = timeit(lambda: a.view(torch.float32).sum(), 10, 3)
    print(f"read-only (sum 4 GiB): {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s", flush=True)


def stem(G, B, cin, iters=5):
    """Inference stem at 256 x 256: conv (stacked samples) + bn_relu_maxpool against the one-kernel stem + bn_act."""
    x = torch.randn(B, cin, 256, 256, device=dev)
    a0 = ops.stem_im2col_f16(x, 7, 7, 2, 3)
    ms_i = timeit(lambda: ops.stem_im2col_f16(x, 7, 7, 2, 3, out=a0), iters)
    print(f"stem im2col B={B} cin={cin}: {ms_i:.3f} ms  {(a0.numel() * 2 + x.numel() * 4) / ms_i / 1e9:.2f} TB/s", flush=True)
    w = (torch.randn(G, 64, a0.shape[1], device=dev) * 0.05).half()
    gamma, beta = torch.ones(64, device=dev), torch.zeros(64, device=dev)
    y = torch.empty(G, B * 16384, 64, device=dev, dtype=torch.float16)
    st = torch.empty(G, B * 128, 64, 2, device=dev)
    ms_c = timeit(lambda: ops.gemm_f16(a0, w, stats=True, shared_a=True, out=y, stats_out=st), iters)
    ss = ops.bn_finalize(st, B * 16384, gamma, beta)
    ms_p = timeit(lambda: ops.bn_relu_maxpool_f16(y.view(G * B, 128, 128, 64), ss, G), iters)
    print(f"stem G={G} B={B} cin={cin}: conv {ms_c:.3f} ms + bn_relu_maxpool {ms_p:.3f} ms = {ms_c + ms_p:.3f} ms", flush=True)
    del y
    ms_f = timeit(lambda: ops.stem_conv_pool_f16(a0, w, B, 128, gamma=gamma), iters)
    yp, st2 = ops.stem_conv_pool_f16(a0, w, B, 128, gamma=gamma)
    ms_a = timeit(lambda: ops.bn_act_f16(yp, ss, G, 64, relu=True, out=yp), iters)
    fl = 2.0 * G * B * 16384 * 64 * 49 * cin
    print(f"   one kernel (conv + statistics + raw max-pool) {ms_f:.3f} ms ({fl / ms_f / 1e9:.0f} TFLOP/s, "
          f"{(yp.numel() * 2 + a0.numel() * 2) / ms_f / 1e9:.2f} TB/s) + bn_act on the pooled tensor {ms_a:.3f} ms = {ms_f + ms_a:.3f} ms", flush=True)


def tail(G, M, N, K, iters=5):
    """Bottleneck tail from the raw conv2 output: bn_act(+colsum) -> second moments + evaluation -> fused conv3, against the
    operand-transform chain (no a2 in HBM)."""
    y = torch.randn(G, M, K, device=dev).half()
    ss = torch.stack([torch.rand(G, K, device=dev) + 0.5, torch.randn(G, K, device=dev) * 0.2], dim=-1).contiguous()
    w = (torch.randn(G, N, K, device=dev) * 0.05).half()
    r = torch.randn(G, M, N, device=dev).half()
    out = torch.empty(G, M, N, device=dev, dtype=torch.float16)
    gamma, beta = torch.ones(N, device=dev), torch.zeros(N, device=dev)
    a = torch.empty_like(y)
    t1 = timeit(lambda: ops.bn_act_f16(y, ss, G, K, relu=True, colsum=True, out=a), iters)
    a, cs = ops.bn_act_f16(y, ss, G, K, relu=True, colsum=True, out=a)
    t2 = timeit(lambda: ops.bn_stats_from_gram(a, cs, w, M, gamma, beta, 1e-5, 0.1), iters)
    ss3 = ops.bn_stats_from_gram(a, cs, w, M, gamma, beta, 1e-5, 0.1)
    t3 = timeit(lambda: ops.gemm_bn_act_f16(a, w, ss3, residual=r, relu=True, out=out), iters)
    print(f"tail G={G} M={M} N={N} K={K}: bn_act {t1:.3f} + moments/eval {t2:.3f} + fused conv3 {t3:.3f} = {t1 + t2 + t3:.3f} ms", flush=True)
    u2 = timeit(lambda: ops.bn_stats_from_gram(y, None, w, M, gamma, beta, 1e-5, 0.1, a_ss=ss), iters)
    u3 = timeit(lambda: ops.gemm_bn_act_f16(y, w, ss3, residual=r, relu=True, out=out, a_ss=ss), iters)
    print(f"   operand transform: moments/eval {u2:.3f} + fused conv3 {u3:.3f} = {u2 + u3:.3f} ms", flush=True)


def layers():
    G, B = 4, 256
    for (M, N, K) in [(B * 4096, 256, 64), (B * 4096, 64, 256), (B * 4096, 64, 64), (B * 1024, 512, 128),
                      (B * 1024, 128, 512), (B * 256, 1024, 256), (B * 256, 256, 1024), (B * 64, 2048, 512),
                      (B * 64, 512, 2048)]:
        gemm(G, M, N, K, 3)
    gemm(G, B * 16384, 64, 152, 3, shared=True)
    for (H, Cin, Cout, k, s) in [(64, 64, 64, 3, 1), (64, 128, 128, 3, 2), (32, 128, 128, 3, 1), (32, 256, 256, 3, 2),
                                 (16, 256, 256, 3, 1), (16, 512, 512, 3, 2), (8, 512, 512, 3, 1), (64, 256, 512, 1, 2)]:
        conv(G, B, H, H, Cin, Cout, k, s, k // 2, 3)


if __name__ == "__main__":
    cmd = sys.argv[1]
    a = [int(x) for x in sys.argv[2:]]
    {"gemm": gemm, "gemm_bn": gemm_bn, "conv": conv, "bnact": bnact, "mcreduce": mcreduce, "kl": kl, "sample": sample, "layers": layers, "hbm": hbm, "hbmwrite": hbmwrite, "gram": gram, "adam": adam, "wgrad": wgrad, "bnbwd": bnbwd, "stem": stem, "tail": tail}[cmd](*a)
