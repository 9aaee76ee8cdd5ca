"""Sharding of reference sequences over GPUs: sequences are independent in vcfeval
(src/com/rtg/vcf/eval/VcfEvalTask.java:131-135 runs one SequenceEvaluator per sequence), so the only multi-GPU
decision is which rank evaluates which sequence.  Longest-processing-time greedy by (estimated) variant count."""
from typing import List, Sequence


def lpt_partition(weights: Sequence[float], n_bins: int) -> List[List[int]]:
    """Assign item i (weight weights[i]) to one of n_bins so that the heaviest bin is as light as possible
    (LPT greedy: heaviest item first, always into the lightest bin).  Deterministic: ties by index."""
    order = sorted(range(len(weights)), key=lambda i: (-weights[i], i))
    bins = [[] for _ in range(n_bins)]
    load = [0.0] * n_bins
    for i in order:
        b = min(range(n_bins), key=lambda k: (load[k], k))
        bins[b].append(i)
        load[b] += weights[i]
    for b in bins:
        b.sort()
    return bins


def imbalance(weights: Sequence[float], bins: List[List[int]]) -> float:
    """max bin load / mean bin load."""
    loads = [sum(weights[i] for i in b) for b in bins]
    mean = sum(loads) / max(1, len(loads))
    return max(loads) / mean if mean > 0 else 1.0
