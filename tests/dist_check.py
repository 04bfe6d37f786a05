"""Multi-GPU check (run under torchrun on the GPU box): MC samples sharded over ranks must reproduce the
single-GPU logits bit for bit, and every rank must end with identical statistics."""
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
from mauv.engine import MCEngine
from mauv.inference.predictors import MCPredictor

bu.dev = f"cuda:{local}"
_, model = bu.build_pair("multimodal")
S = 7
img, bathy, sss, _ = O.synthetic_batch(4, size=64)
xs = [t.cuda() for t in (img, bathy, sss)]
pred = MCPredictor(model, S)
out = pred.predict_device(xs, seed=99, sample0=0)
single = MCEngine(model).forward_mc(xs, S, seed=99, sample0=0)          # all samples on this rank
ok = torch.equal(out["logits"], single)
# end-to-end staging: every rank is handed the whole pinned host batch, uploads only its 1/N slice of the rows and the slices
# are all-gathered over NVLink (mauv.inference.predictors.stage_batch) - same results as device-resident inputs
from mauv.inference.predictors import predict_stream
host = [t.pin_memory() for t in (img, bathy, sss)]
c0 = pred.engine._sample_cursor
res = list(predict_stream(pred, [host, host]))
ref0 = pred.predict_device(xs, sample0=c0)
ref1 = pred.predict_device(xs, sample0=c0 + S)
ok = ok and torch.equal(res[0]["predicted_class"], ref0["argmax_prob"].cpu()) and \
    torch.equal(res[0]["aleatoric_uncertainty"], ref0["aleatoric"].cpu()) and \
    torch.equal(res[1]["aleatoric_uncertainty"], ref1["aleatoric"].cpu())
gathered = [torch.empty_like(out["mean_prob"]) for _ in range(world)]
torch.distributed.all_gather(gathered, out["mean_prob"])
same = all(torch.equal(g, gathered[0]) for g in gathered)
print(f"rank {rank}/{world}: sharded == single-GPU logits: {ok}; statistics identical on all ranks: {same}", flush=True)
torch.distributed.barrier()
torch.distributed.destroy_process_group()
sys.exit(0 if ok and same else 1)
