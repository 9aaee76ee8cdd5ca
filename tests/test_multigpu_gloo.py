"""The N>1 path on the CPU: two gloo ranks shard the sequences (LPT), evaluate their shards (through the test-only
emulation of the engine -- there is no GPU here), merge the summary counters with an all-reduce and gather the ROC
tuples; the merged result must equal the single-process result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build():
    from rtg_tools_b200 import synth
    seqs, quals = synth.make_pair(6_000_000, n_chrom=6, return_qual=True)
    return seqs, quals


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import CONFIGS, emul_library
    from rtg_tools_b200 import multigpu
    seqs, quals = _build()
    weights = [s.baseline.n + s.calls.n for s in seqs]
    mine = multigpu.shard_sequences(weights, rank, world)
    eng = emul_library()
    res = eng.submit(CONFIGS["default"], [seqs[i] for i in mine])
    counts = multigpu.merge_counts(multigpu.summary_counts(res))
    blocks = [multigpu.roc_tuples([r], [quals[i]]) for r, i in zip(res, mine)]
    roc = multigpu.gather_roc(blocks, mine, len(seqs))
    if rank == 0:
        np.save(os.path.join(out_dir, "counts.npy"), counts)
        np.save(os.path.join(out_dir, "roc.npy"), roc)
    np.save(os.path.join(out_dir, "mine%d.npy" % rank), np.asarray(mine))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_evaluation_merges_to_the_single_process_result(tmp_path, oracle, world):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import CONFIGS
    from rtg_tools_b200 import multigpu
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    seqs, quals = _build()
    want = oracle.submit(CONFIGS["default"], seqs)
    want_counts = multigpu.summary_counts(want)
    want_roc = multigpu.roc_tuples(want, quals)
    got_counts = np.load(tmp_path / "counts.npy")
    got_roc = np.load(tmp_path / "roc.npy")
    assert np.array_equal(got_counts, want_counts)
    assert got_roc.shape == want_roc.shape
    assert np.array_equal(got_roc.view(np.int64), want_roc.view(np.int64)), "ROC tuples must arrive in sequence/record order"
    # every sequence evaluated exactly once, shards balanced
    owned = sorted(int(x) for r in range(world) for x in np.load(tmp_path / ("mine%d.npy" % r)))
    assert owned == list(range(len(seqs)))
    # K4 on the merged tuples equals K4 on the single-process tuples (same arrival order -> same FP64 sums)
    a, na = oracle.roc(got_roc[:, 0], got_roc[:, 1], got_roc[:, 2], got_roc[:, 3], None, 1)
    b, nb = oracle.roc(want_roc[:, 0], want_roc[:, 1], want_roc[:, 2], want_roc[:, 3], None, 1)
    assert na == nb
    for x, y in zip(a[0], b[0]):
        assert np.array_equal(x.view(np.int64), y.view(np.int64))


def test_lpt_partition_balances_by_variant_count():
    from rtg_tools_b200 import partition, synth
    lens = synth.genome_lengths(3_000_000_000, 24)
    for world in (2, 4, 8):
        bins = partition.lpt_partition(lens, world)
        assert sorted(i for b in bins for i in b) == list(range(24))
        assert partition.imbalance(lens, bins) < 1.12


def test_merge_is_a_no_op_without_a_process_group():
    from rtg_tools_b200 import multigpu
    c = np.arange(8, dtype=np.int64)
    assert np.array_equal(multigpu.merge_counts(c), c)
    rows = multigpu.gather_roc([np.ones((2, 4)), np.zeros((1, 4))], [1, 0], 2)
    assert rows.shape == (3, 4) and rows[0, 0] == 0.0


@pytest.mark.parametrize("name", ["annotate", "updel2", "updel4", "split", "tricky", "small"])
def test_roc_tuples_follow_the_reference_synchronizer_rules(oracle, fixtures, name):
    """The merge's status -> (tp, fp, raw) mapping against the writers' restatement in oracle/ref_host.py (which the golden
    fixtures pin): two-pass scenarios exercise ALLELE_MATCH -> FP and weight-0 matches, `split` the OUTSIDE_EVAL rules."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_host
    from rtg_tools_b200 import multigpu
    sc = fixtures[name]
    rr = ref_host.run_vcfeval(oracle, sc["template_fa"], sc["baseline_vcf"], sc["calls_vcf"], sc["args"])
    scores = [np.array([ref_host.score_of(v["rec"], rr.params.score_field, rr.csample) for v in cl], np.float64)
              for _, cl in rr.loaded]
    got = multigpu.roc_tuples(rr.results, scores)
    want = np.array([rr.roc_inputs["score"], rr.roc_inputs["tp"], rr.roc_inputs["fp"], rr.roc_inputs["raw"]], np.float64).T.reshape(-1, 4)
    assert got.shape == want.shape
    assert np.array_equal(np.nan_to_num(got, nan=-1.0), np.nan_to_num(want, nan=-1.0))
    counts = multigpu.summary_counts(rr.results)
    assert counts[0] == rr.base_tp and counts[0] + counts[1] == rr.base_total
    assert counts[2] + counts[3] == got.shape[0]
