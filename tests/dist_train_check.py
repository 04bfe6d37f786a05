"""2-GPU data-parallel ELBO step (run under torchrun on the GPU box): DistributedDataParallel over the drop-in model.
Each rank takes half of the minibatch (per-replica BatchNorm statistics, as the reference's nn.DataParallel:
utils/device.py:19); after backward every rank must hold the mean of the two shard gradients (NCCL all-reduce)."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "multimodal-auv_b200", ROOT / "oracle", ROOT / "tests"):
    sys.path.insert(0, str(p))
import torch

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
import bnn_oracle as O
import gpu_bringup as bu
import mauv.bayesian as MB

bu.dev = f"cuda:{local}"
_, model = bu.build_pair("multimodal")
B, S = 4, 2
img, bathy, sss, labels = O.synthetic_batch(B, size=64)
lo, hi = rank * B // world, (rank + 1) * B // world


def step(m, sl):
    for l in [l for _, l in MB.bayesian_layers(m)]:
        l._calls = 0                                   # same Philox sample ids on every rank and every call
    xs = [t[sl].cuda() for t in (img, bathy, sss)]
    out = torch.mean(torch.stack([m(*xs) for _ in range(S)]), dim=0)
    loss = torch.nn.functional.cross_entropy(out, labels[sl].cuda()) + MB.get_kl_loss(m) / B * O.kl_weight(1, 20)
    m.zero_grad(set_to_none=True)
    loss.backward()
    return {n: p.grad.detach().clone() for n, p in m.named_parameters()}


MB.manual_seed(1)
# reference: the two shard gradients computed locally, without communication, then averaged
g0 = step(model, slice(0, B // 2))
g1 = step(model, slice(B // 2, B))
want = {n: 0.5 * (g0[n] + g1[n]) for n in g0}
# S model calls precede one backward: a buffer broadcast at call 2 would overwrite BN running stats autograd saved at call 1
ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], broadcast_buffers=False)
got = step(ddp, slice(lo, hi))
worst = 0.0
for n, g in got.items():
    ref = want[n.replace("module.", "")]
    worst = max(worst, ((g - ref).abs().max() / (ref.abs().max() + 1e-30)).item())
ok = worst < 2e-2          # BN running statistics differ between the runs; gradients only through fp16 transport noise
print(f"rank {rank}/{world}: DDP gradient == mean of shard gradients, worst rel-to-max deviation {worst:.2e} -> {ok}", flush=True)

# ---- the S-batched engine: shard gradients averaged by ONE all-reduce over the flat gradient buffer
from mauv.train_engine import TrainEngine
eng = TrainEngine(model)
eng.flatten_grads()
kl_scale = O.kl_weight(1, 20) / B


def estep(sl, reduce):
    eng.zero_grad()
    eng.step([t[sl].cuda() for t in (img, bathy, sss)], labels[sl], S, kl_scale, sample0=0)
    if reduce:
        eng.allreduce_grads()
    return {n: p.grad.detach().clone() for n, p in model.named_parameters()}


e0, e1 = estep(slice(0, B // 2), False), estep(slice(B // 2, B), False)
got = estep(slice(lo, hi), True)
worst_e = 0.0
for n, g in got.items():
    ref = 0.5 * (e0[n] + e1[n])
    worst_e = max(worst_e, ((g - ref).abs().max() / (ref.abs().max() + 1e-30)).item())
ok_e = worst_e < 1e-5
print(f"rank {rank}/{world}: engine all-reduced gradient == mean of shard gradients, worst deviation {worst_e:.2e} -> {ok_e}", flush=True)
torch.distributed.barrier()
torch.distributed.destroy_process_group()
sys.exit(0 if (ok and ok_e) else 1)
