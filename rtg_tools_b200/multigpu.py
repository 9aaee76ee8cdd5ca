"""Multi-GPU plumbing of the vcfeval path: one process per GPU, sequences sharded over ranks, one exchange at the end.

Reference sequences are independent units of work in vcfeval (src/com/rtg/vcf/eval/VcfEvalTask.java:131-135 runs one
SequenceEvaluator per sequence), so no collective is needed on the data path.  The only exchange is the final merge
that the reference does implicitly by sharing one RocContainer / EvalSynchronizer between its worker threads
(RocContainer.java:204-258, EvalSynchronizer.addPhasingCounts):
  * int64 all-reduce of the summary / phasing counters;
  * gather of the per-call ROC tuples onto the rank that runs K4, in (sequence order, record order) -- RocPoint.add
    accumulates doubles in arrival order (RocContainer.java:251-258), so the order is part of the result.
The functions take any torch.distributed backend: NCCL over NVLink on the B200 box, gloo in the CPU tests.
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import capi, partition


def shard_sequences(weights: Sequence[float], rank: int, world: int) -> List[int]:
    """Indices of the sequences rank `rank` evaluates: longest-processing-time greedy by variant count (or length)."""
    return partition.lpt_partition(weights, world)[rank]


def _device(backend_device: Optional[torch.device]) -> torch.device:
    if backend_device is not None:
        return backend_device
    if not dist.is_initialized():
        return torch.device("cpu")
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def merge_counts(counts: np.ndarray, device: Optional[torch.device] = None) -> np.ndarray:
    """Sum of the per-rank int64 counters (TP / FN / FP totals, phasing counts, skipped ...)."""
    counts = np.ascontiguousarray(counts, np.int64)
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return counts
    t = torch.from_numpy(counts).to(_device(device))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def call_classes(status: np.ndarray, weight: np.ndarray):
    """Boolean masks (tp, fp_ca, fp, outside) of the calls of one sequence, split output mode
    (SplitEvalSynchronizer.handleKnownCall :112-134): GT_MATCH with weight > 0 -> TP (a self-cancelling call with weight 0
    is dropped), else ALLELE_MATCH with weight > 0 -> FP_CA, else NO_MATCH -> FP unless OUTSIDE_EVAL."""
    st = np.asarray(status, np.uint8)
    w = np.asarray(weight, np.float64)
    gt = (st & capi.STATUS_GT_MATCH) != 0
    am = ((st & capi.STATUS_ALLELE_MATCH) != 0) & ~gt
    nm = ((st & capi.STATUS_NO_MATCH) != 0) & ~gt & ~am
    out = (st & capi.STATUS_OUTSIDE_EVAL) != 0
    tp = gt & (w > 0)
    fp_ca = am & (w > 0)
    fp = nm & ~out
    return tp, fp_ca, fp, nm & out


def baseline_classes(status: np.ndarray):
    """(tp, fn_ca, fn) masks of the baseline variants of one sequence (SplitEvalSynchronizer.handleKnownBaseline :136-150):
    OUTSIDE_EVAL variants are not counted at all."""
    st = np.asarray(status, np.uint8)
    inside = (st & capi.STATUS_OUTSIDE_EVAL) == 0
    gt = (st & capi.STATUS_GT_MATCH) != 0
    am = ((st & capi.STATUS_ALLELE_MATCH) != 0) & ~gt
    nm = ((st & capi.STATUS_NO_MATCH) != 0) & ~gt & ~am
    return gt & inside, am & inside, nm & inside


def summary_counts(results: List[capi.SequenceResult]) -> np.ndarray:
    """[baseline TP, baseline FN (incl. FN_CA), call TP, call FP (incl. FP_CA), skipped, mis-phasings, correct phasings,
    unphaseable] with the reference's split-mode rules (see call_classes / baseline_classes)."""
    tot = np.zeros(8, np.int64)
    for r in results:
        btp, bca, bfn = baseline_classes(r.baseline.status)
        ctp, cca, cfp, _ = call_classes(r.calls.status, r.calls.weight)
        tot[0] += int(np.count_nonzero(btp))
        tot[1] += int(np.count_nonzero(bca)) + int(np.count_nonzero(bfn))
        tot[2] += int(np.count_nonzero(ctp))
        tot[3] += int(np.count_nonzero(cca)) + int(np.count_nonzero(cfp))
        tot[4] += int(np.count_nonzero(r.baseline.status & np.uint8(capi.STATUS_SKIPPED)))
        tot[4] += int(np.count_nonzero(r.calls.status & np.uint8(capi.STATUS_SKIPPED)))
        tot[5:8] += np.asarray(r.phasing, np.int64)
    return tot


def roc_tuples(results: List[capi.SequenceResult], scores: List[np.ndarray]) -> np.ndarray:
    """(n, 4) float64 rows (score, tpWeight, fpWeight, tpRaw) of the calls of this rank's sequences, in sequence then
    record order, as WithRocsEvalSynchronizer.addToROCContainer :132-141 feeds the default ROC: TP -> (w, 0, 1), an allele
    match (two-pass) -> (0, 1, 0), FP -> (0, 1, 0); weight-0 matches and OUTSIDE_EVAL calls contribute nothing."""
    rows = []
    for r, sc in zip(results, scores):
        tp, fp_ca, fp, _ = call_classes(r.calls.status, r.calls.weight)
        keep = tp | fp_ca | fp
        t = np.zeros((int(keep.sum()), 4), np.float64)
        t[:, 0] = np.asarray(sc, np.float64)[keep]
        t[:, 1] = np.where(tp, r.calls.weight, 0.0)[keep]
        t[:, 2] = (fp_ca | fp)[keep].astype(np.float64)
        t[:, 3] = tp[keep].astype(np.float64)
        rows.append(t)
    return np.concatenate(rows) if rows else np.zeros((0, 4), np.float64)


def gather_roc(local_rows: List[np.ndarray], local_seq_ids: List[int], n_seq_total: int,
               device: Optional[torch.device] = None) -> Optional[np.ndarray]:
    """Gathers the per-sequence ROC tuple blocks of every rank; rank 0 returns them concatenated in global sequence
    order (other ranks return None).  local_rows[i] is the (n_i, 4) block of global sequence local_seq_ids[i]."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        order = np.argsort(np.asarray(local_seq_ids, np.int64), kind="stable")
        return np.concatenate([local_rows[i] for i in order]) if len(order) else np.zeros((0, 4), np.float64)
    dev = _device(device)
    # per-rank table: rows per global sequence (0 for sequences of other ranks)
    sizes = torch.zeros(n_seq_total, dtype=torch.int64)
    for sid, rows in zip(local_seq_ids, local_rows):
        sizes[sid] = rows.shape[0]
    all_sizes = [torch.zeros(n_seq_total, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes.to(dev))
    all_sizes = torch.stack([s.cpu() for s in all_sizes])          # (world, n_seq)
    per_rank = all_sizes.sum(dim=1)
    cap = int(per_rank.max().item())
    # this rank's blocks in ascending global sequence id, padded to the largest rank
    order = np.argsort(np.asarray(local_seq_ids, np.int64), kind="stable")
    mine = np.concatenate([local_rows[i] for i in order]) if len(order) else np.zeros((0, 4), np.float64)
    buf = torch.zeros((cap, 4), dtype=torch.float64)
    buf[: mine.shape[0]] = torch.from_numpy(np.ascontiguousarray(mine))
    gathered = [torch.zeros((cap, 4), dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(gathered, buf.to(dev))
    if rank != 0:
        return None
    gathered = [g.cpu().numpy() for g in gathered]
    cursor = [0] * world
    out = []
    for sid in range(n_seq_total):
        for r in range(world):
            n = int(all_sizes[r, sid].item())
            if n:
                out.append(gathered[r][cursor[r]:cursor[r] + n])
                cursor[r] += n
    return np.concatenate(out) if out else np.zeros((0, 4), np.float64)
