"""CPU tests of the host batcher's planning stage (length binning, batch cutting, LPT sharding) and of the
multi-process shard/gather plumbing (torch.distributed, gloo, world_size 2)."""
import os
import subprocess
import sys

import numpy as np

from genomics_gpu_b200.aligner import plan_batches

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lens(seed=0, n=20000):
    rng = np.random.default_rng(seed)
    q = np.exp(rng.uniform(np.log(100), np.log(10000), size=n)).astype(np.uint32)
    t = (q * rng.uniform(1.0, 1.3, size=n)).astype(np.uint32)
    return q, t


def test_plan_is_a_permutation_and_respects_limits():
    q, t = lens()
    p = plan_batches(q, t, n_gpus=4, max_batch_pairs=1000, max_batch_cells=1 << 33)
    assert np.array_equal(np.sort(p["order"]), np.arange(q.size))
    st = p["batch_start"]
    assert st[0] == 0 and st[-1] == q.size and np.all(np.diff(st.astype(np.int64)) > 0)
    cells = q.astype(np.uint64) * t
    for b in range(len(st) - 1):
        idx = p["order"][st[b]:st[b + 1]]
        assert idx.size <= 1000
        assert int(cells[idx].sum()) == int(p["batch_cells"][b])
        assert idx.size == 1 or int(cells[idx].sum()) <= (1 << 33) + int(cells[idx].max())
        q8 = (q[idx] + 7) // 8 * 8
        assert (q8 <= 352).all() or (q8 > 352).all(), "fast-path and general pairs must not share a batch"


def test_length_binning_sorts_by_shape():
    q, t = lens(1)
    p = plan_batches(q, t, n_gpus=1, max_batch_pairs=512)
    qs = ((q[p["order"]] + 7) // 8).astype(np.int64)
    assert np.all(np.diff(qs) >= 0), "pairs must be ordered by padded query length class"
    keep = plan_batches(q, t, n_gpus=1, max_batch_pairs=512, keep_input_order=True)
    assert np.array_equal(keep["order"], np.arange(q.size))


def test_lpt_balances_cells_over_gpus():
    q, t = lens(2)
    for g in (2, 4, 8):
        p = plan_batches(q, t, n_gpus=g, max_batch_pairs=256)
        load = np.zeros(g)
        for b, dev in enumerate(p["batch_gpu"]):
            load[dev] += float(p["batch_cells"][b])
        assert load.min() > 0
        # LPT guarantee: max load <= mean + largest batch
        assert load.max() <= load.mean() + float(p["batch_cells"].max()) + 1
        assert load.max() / load.mean() < 1.05


def test_empty_and_single():
    p = plan_batches(np.array([], np.uint32), np.array([], np.uint32), n_gpus=2)
    assert p["order"].size == 0 and p["batch_cells"].size == 0
    p = plan_batches(np.array([150], np.uint32), np.array([600], np.uint32), n_gpus=8)
    assert p["batch_cells"].tolist() == [90000] and p["batch_gpu"].tolist() == [0]


def test_multiprocess_shard_and_gather_gloo():
    """world_size 2 over gloo: rank 0 plans, both ranks 'align' their shard (with the CPU oracle standing in for the GPU),
    results are gathered on rank 0 in the original order (genomics_gpu_b200/sharding.py)."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", script], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "GLOO_SHARD_GATHER_OK" in p.stdout
