"""Drop-in for the reference's inference/predictors.py: `multimodal_predict_and_save` keeps its signature,
CSV header/rows and console output; the S Monte-Carlo passes run S-batched through engine.MCEngine and
the uncertainty statistics in one CUDA kernel (K5) instead of ~10 ATen launches + 3*B `.item()` syncs.

With torch.distributed initialised (one process per GPU) the MC samples are block-partitioned over the
ranks (disjoint Philox sample ids), the per-rank logits are all-gathered over NCCL and every rank reduces
the full [S, B, C] stack, so results are identical to a single-GPU run.
"""
from __future__ import annotations

import csv
import logging
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .. import ops
from ..engine import MCEngine


def shard_samples(S: int, world: int, rank: int) -> Tuple[int, int]:
    """Block partition of sample ids [0, S) -> [lo, hi) for `rank` (first S % world ranks get one extra)."""
    base, rem = divmod(S, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_sample_blocks(local: Optional[torch.Tensor], S: int, B: int, C: int, world: int, device) -> torch.Tensor:
    """All ranks contribute logits of their sample block ([hi-lo, B, C], possibly empty) and receive the full
    [S, B, C] stack in sample order. The only collective on the inference path: S*B*C*4 bytes."""
    per = (S + world - 1) // world
    pad = torch.zeros((per, B, C), dtype=torch.float32, device=device)
    if local is not None and local.shape[0] > 0:
        pad[: local.shape[0]] = local
    gathered = torch.empty((world * per, B, C), dtype=torch.float32, device=device)   # concatenated along dim 0
    torch.distributed.all_gather_into_tensor(gathered, pad)
    gathered = gathered.view(world, per, B, C)
    parts = []
    for r in range(world):
        rlo, rhi = shard_samples(S, world, r)
        parts.append(gathered[r, : rhi - rlo])
    return torch.cat(parts, dim=0).contiguous()


class MCPredictor:
    """H2D -> S-batched MC forward -> MC statistics -> one D2H, for one batch."""

    def __init__(self, model: nn.Module, num_mc_samples: int, group: Optional[int] = None, eps_entropy: float = 1e-7,
                 use_graph="auto"):
        self.engine = MCEngine(model, max_group=group)      # group None: as many samples per walk as memory allows
        # CUDA graph of this rank's S-pass forward (all sample groups): one replay instead of ~2.4 k Python-driven
        # launches per sample group. Python enqueues ~57 us per launch (414 ms per cfg2 step), which is hidden behind
        # the GPU at B=256 / G=10 (93 us of GPU work per launch) but bounds small batches (cfg1) and the 8-GPU
        # sample-sharded case (G=4); a graph of thousands of nodes on the other hand costs ~60 ms of launch latency
        # that is exposed when every batch ends with a D2H sync. "auto" picks the graph when the estimated GPU time per
        # launch is below the Python enqueue cost. Keyed by input shapes / sample range / seed; inputs are copied
        # into static buffers. Injected-eps (validation) calls always run eagerly.
        self.use_graph = use_graph
        self._graphs = {}
        self.S = int(num_mc_samples)
        self.eps_entropy = eps_entropy
        self.device = self.engine.device
        self._staging = None
        self._slots = None
        self.dist = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.world = torch.distributed.get_world_size() if self.dist else 1
        self.rank = torch.distributed.get_rank() if self.dist else 0

    @torch.no_grad()
    def mc_logits(self, inputs: Sequence[torch.Tensor], eps: Optional[dict] = None,
                  seed: Optional[int] = None, sample0: Optional[int] = None) -> torch.Tensor:
        """[S, B, C] logits; under torch.distributed each rank computes its block and all ranks gather.
        sample0: first Philox sample id of this batch's S draws. None (production) takes a fresh block of S ids from the
        model's cursor on every call - identical on all ranks, which advance in lockstep - so every batch is evaluated
        with its own Monte-Carlo draws, as in the reference; with injected eps the ids index the eps tensors (base 0)."""
        from .. import engine as _engine
        if sample0 is None:
            sample0 = 0 if (eps is not None or _engine.DEBUG_EPS is not None) else self.engine.take_samples(self.S)
        lo, hi = shard_samples(self.S, self.world, self.rank)
        if hi <= lo:
            local = None
        elif eps is None and self._want_graph(inputs[0].shape[0], hi - lo):
            local = self._forward_graphed(inputs, sample0, lo, hi, seed)
        else:
            local = self.engine.forward_mc(inputs, hi - lo, sample0=sample0 + lo, eps=eps, seed=seed)
        if self.world == 1:
            return local
        C = local.shape[-1] if local is not None else self._num_classes()
        return gather_sample_blocks(local, self.S, inputs[0].shape[0], C, self.world, self.device)

    def staging_stream(self) -> torch.cuda.Stream:
        if self._staging is None:
            self._staging = torch.cuda.Stream(self.device)
        return self._staging

    def _staging_slot(self, host_inputs) -> dict:
        """Two persistent sets of device input buffers (double buffering), re-created only when the batch geometry changes."""
        key = tuple((tuple(x.shape), x.dtype) for x in host_inputs)
        if self._slots is None or self._slots[0] != key:
            bufs = []
            for _ in range(2):
                one = []
                for x in host_inputs:
                    rows = x.shape[0] if self.world == 1 else batch_slice(x.shape[0], self.world, 0)[2] * self.world
                    one.append(torch.empty((rows, *x.shape[1:]), dtype=x.dtype, device=self.device))
                bufs.append({"bufs": one, "done": None})
            self._slots = (key, bufs, 0)
        key, bufs, nxt = self._slots
        self._slots = (key, bufs, nxt ^ 1)
        return bufs[nxt]

    def _want_graph(self, B: int, S_local: int) -> bool:
        if self.use_graph == "auto":
            G = min(self.engine.max_group or S_local, S_local)
            return 0.036 * G * B < 60.0          # measured: 93 us GPU time per launch at G=10, B=256; 57 us to enqueue
        return bool(self.use_graph)

    def _forward_graphed(self, inputs, sample0, lo, hi, seed):
        from .. import engine as _engine
        from ..bayesian import current_seed
        if _engine.DEBUG_EPS is not None:
            return self.engine.forward_mc(inputs, hi - lo, sample0=sample0 + lo, seed=seed)
        seed = current_seed() if seed is None else seed
        # the graph is recorded for sample ids lo .. hi-1 PLUS a device-resident base word (ops.sample_base): the batch's
        # sample0 is written into that word before every replay, so one graph serves every batch with fresh draws
        # (the first parameter's address is part of the key: FusedAdam / load_state_dict may re-home the weights)
        key = (tuple(tuple(x.shape) for x in inputs), lo, hi, seed, self.engine.precision,
               next(self.engine.model.parameters()).data_ptr())
        entry = self._graphs.get(key)
        if entry is None:
            static_in = [torch.empty(x.shape, dtype=torch.float32, device=self.device) for x in inputs]
            for d, x in zip(static_in, inputs):
                d.copy_(x)
            base = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.engine.forward_mc(static_in, hi - lo, sample0=lo, seed=seed)       # eager warm-up (lazy inits)
            torch.cuda.synchronize(self.device)
            n0 = ops.launch_count
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph), ops.sample_base(base):
                static_out = self.engine.forward_mc(static_in, hi - lo, sample0=lo, seed=seed)
            entry = (graph, static_in, static_out, ops.launch_count - n0, base)
            self._graphs[key] = entry
        graph, static_in, static_out, n_launch, base = entry
        for d, x in zip(static_in, inputs):
            d.copy_(x, non_blocking=True)
        base.fill_(((sample0 & 0xFFFFFFFF) ^ 0x80000000) - 0x80000000)      # uint32 bit pattern in an int32 word
        graph.replay()
        ops.launch_count += n_launch
        return static_out.clone()      # 4*S*B*C bytes; the static buffer is overwritten by the next replay

    def _num_classes(self) -> int:
        m = self.engine.model
        return m.fc2.out_features if hasattr(m, "fc2") else m.model.fc.out_features

    @torch.no_grad()
    def predict_device(self, inputs: Sequence[torch.Tensor], eps: Optional[dict] = None,
                       seed: Optional[int] = None, sample0: Optional[int] = None) -> Dict[str, torch.Tensor]:
        logits = self.mc_logits(inputs, eps, seed, sample0)
        out = ops.mc_reduce(logits, self.eps_entropy)
        out["logits"] = logits
        return out

    @torch.no_grad()
    def predict_batch(self, host_inputs: Sequence[torch.Tensor], eps: Optional[dict] = None,
                      seed: Optional[int] = None, sample0: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """Host tensors in, host results out: (class [B] int64, predictive unc. [B], aleatoric [B], MI [B])."""
        st = stage_batch(self, host_inputs)                                   # 1/N slice per rank + NVLink all-gather
        torch.cuda.current_stream(self.device).wait_event(st.event)
        o = self.predict_device(st.tensors, eps, seed, sample0)
        st.release(self.device)
        return _unpack_results(_pack_results(o).cpu())                       # one [B, 5] D2H


def _pack_results(o: Dict[str, torch.Tensor]) -> torch.Tensor:
    return torch.stack([o["argmax_prob"].to(torch.float32), o["var_mean"], o["aleatoric"],
                        o["pred_entropy"], o["mutual_info"]], dim=1)


def _unpack_results(host: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"predicted_class": host[:, 0].to(torch.int64), "predictive_uncertainty": host[:, 1],
            "aleatoric_uncertainty": host[:, 2], "pred_entropy": host[:, 3], "mutual_info": host[:, 4]}


def batch_slice(B: int, world: int, rank: int) -> Tuple[int, int, int]:
    """Row range [lo, hi) of a batch that `rank` uploads, and the common (padded) chunk length of the gather."""
    chunk = (B + world - 1) // world
    lo = min(rank * chunk, B)
    return lo, min(lo + chunk, B), chunk


class StagedBatch:
    """Device-resident inputs of one batch, the event that marks them complete on the staging stream, and the staging slot
    they occupy (released - for the staging stream - by `done`, recorded on the compute stream after the batch's forward)."""

    def __init__(self, tensors, event, h2d_bytes, slot=None):
        self.tensors, self.event, self.h2d_bytes, self.slot = tensors, event, h2d_bytes, slot

    def release(self, device) -> None:
        if self.slot is not None:
            self.slot["done"] = torch.cuda.Event()
            self.slot["done"].record(torch.cuda.current_stream(device))


def stage_batch(predictor: "MCPredictor", host_inputs: Sequence[torch.Tensor]) -> StagedBatch:
    """Input staging (the reference's loaders hand every process the whole pinned batch,
    data/loaders.py:48-51): enqueue, on the predictor's staging stream, the host->device copy of one batch into one of two
    persistent device slots (no allocator traffic, no implicit synchronisation). With N ranks every rank uploads only ITS
    1/N slice of the rows and the slices are all-gathered over NVLink, so the PCIe / host-memory traffic per batch is B rows
    in total instead of N x B. Nothing here waits on the host: the caller overlaps it with the previous batch's compute and
    makes the compute stream wait on `.event`."""
    dev, world, rank = predictor.device, predictor.world, predictor.rank
    side = predictor.staging_stream()
    slot = predictor._staging_slot(host_inputs)
    out, nbytes = [], 0
    with torch.cuda.stream(side):
        if slot["done"] is not None:
            side.wait_event(slot["done"])          # the batch that used this slot two calls ago has been consumed
        for x, full in zip(host_inputs, slot["bufs"]):
            if x.is_cuda:
                out.append(x)
                continue
            B = x.shape[0]
            if world == 1:
                full.copy_(x, non_blocking=True)
                nbytes += x.numel() * x.element_size()
                out.append(full)
            else:
                lo, hi, chunk = batch_slice(B, world, rank)
                mine = full[rank * chunk:(rank + 1) * chunk]
                if hi > lo:
                    mine[: hi - lo].copy_(x[lo:hi], non_blocking=True)
                    nbytes += (hi - lo) * x[0].numel() * x.element_size()
                torch.distributed.all_gather_into_tensor(full, mine)        # in place: rank r's chunk is rows [r*chunk, ..)
                out.append(full[:B])
        ev = torch.cuda.Event()
        ev.record(side)
    return StagedBatch(out, ev, nbytes, slot)


def predict_stream(predictor: "MCPredictor", host_batches):
    """predict_batch over an iterable of host batches, one result dict (host tensors) per batch. The staging of batch i+1
    (H2D of this rank's slice + NVLink all-gather, `stage_batch`) is enqueued on a side stream BEFORE batch i is computed,
    so it overlaps the compute; the results of batch i are read back with one [B, 5] D2H."""
    it = iter(host_batches)
    try:
        nxt = stage_batch(predictor, next(it))
    except StopIteration:
        return
    while nxt is not None:
        cur = nxt
        hb = next(it, None)
        nxt = stage_batch(predictor, hb) if hb is not None else None
        torch.cuda.current_stream(predictor.device).wait_event(cur.event)
        o = predictor.predict_device(cur.tensors)
        cur.release(predictor.device)
        yield _unpack_results(_pack_results(o).cpu())


def multimodal_predict_and_save(multimodal_model: nn.Module, dataloader, device: torch.device, csv_path: str,
                                num_mc_samples: int = 10, sss_patch_type: Optional[str] = "",
                                channel_patch_type: Optional[str] = "", model_type: str = "multimodal"):
    """Same contract as reference inference/predictors.py:9-97 (CSV: Image Name, Predicted Class,
    Predictive Uncertainty = var_s(p).mean_c, Aleatoric Uncertainty = mean_s H[p_s], class = argmax mean_s p)."""
    multimodal_model.train()  # BN batch statistics per MC pass, as the reference (predictors.py:27)
    if isinstance(multimodal_model, (nn.parallel.DistributedDataParallel, nn.DataParallel)):
        multimodal_model = multimodal_model.module
    predictor = MCPredictor(multimodal_model, num_mc_samples)
    logging.info(f"CSV will be saved to: {csv_path}")
    with open(csv_path, mode="w", newline="") as csvfile:
        csv_writer = csv.writer(csvfile)
        header = ["Image Name", "Predicted Class", "Predictive Uncertainty", "Aleatoric Uncertainty"]
        csv_writer.writerow(header)
        logging.info(f"CSV Header written: {header}")
        logging.info(f"Length of the dataloader: {len(dataloader)}")
        names = []

        def host_batches():
            for inputs, patch_30_bathy, patch_30_sss, image_name in dataloader:
                names.append((inputs.size(0), image_name))
                yield (inputs, patch_30_bathy, patch_30_sss)

        for batch_idx, res in enumerate(predict_stream(predictor, host_batches())):
            logging.info(f"\n--- Processing Batch {batch_idx + 1} ---")
            n_rows, image_name = names[batch_idx]
            print(f"Predictive Uncertainty: {res['predictive_uncertainty'].numpy()}")
            print(f"Aleatoric Uncertainty: {res['aleatoric_uncertainty'].numpy()}")
            print(f"Predicted Classes: {res['predicted_class'].numpy()}")
            cls = res["predicted_class"].tolist()
            pu = res["predictive_uncertainty"].tolist()
            au = res["aleatoric_uncertainty"].tolist()
            for i in range(n_rows):
                name = image_name[i] if isinstance(image_name, (list, tuple)) else image_name
                csv_writer.writerow([name, cls[i], pu[i], au[i]])
    logging.info("Completed: multimodal_predict_and_save")
