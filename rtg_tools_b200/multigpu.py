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


def _flag_counts(status: np.ndarray, flags) -> List[int]:
    """How many status bytes carry each of the single-bit `flags` (two passes per flag, no boolean temporaries)."""
    return [int(np.count_nonzero(status & np.uint8(f))) for f in flags]


def summary_counts(results: List[capi.SequenceResult]) -> np.ndarray:
    """[baseline TP, baseline FN, call TP, call FP, skipped, mis-phasings, correct phasings, unphaseable]."""
    tot = np.zeros(8, np.int64)
    flags = (capi.STATUS_GT_MATCH, capi.STATUS_NO_MATCH, capi.STATUS_SKIPPED)
    for r in results:
        b = _flag_counts(r.baseline.status, flags)
        c = _flag_counts(r.calls.status, flags)
        tot[0] += b[0]
        tot[1] += b[1]
        tot[2] += c[0]
        tot[3] += c[1]
        tot[4] += b[2] + c[2]
        tot[5:8] += np.asarray(r.phasing, np.int64)
    return tot


def roc_tuples(results: List[capi.SequenceResult], scores: List[np.ndarray]) -> np.ndarray:
    """(n, 4) float64 rows (score, tpWeight, fpWeight, tpRaw) of the calls of this rank's sequences, in sequence then
    record order: TP -> (w, 0, 1), FP -> (0, 1, 0)  (WithRocsEvalSynchronizer.java:132-141)."""
    rows = []
    for r, sc in zip(results, scores):
        st = r.calls.status
        tp = (st & capi.STATUS_GT_MATCH) != 0
        fp = (st & capi.STATUS_NO_MATCH) != 0
        keep = tp | fp
        t = np.zeros((int(keep.sum()), 4), np.float64)
        t[:, 0] = np.asarray(sc, np.float64)[keep]
        t[:, 1] = np.where(tp, r.calls.weight, 0.0)[keep]
        t[:, 2] = fp[keep].astype(np.float64)
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
