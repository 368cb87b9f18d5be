"""The N > 1 path on CPU: world_size-2 gloo process group, each rank computes its cell-balanced shard (the compute
function is injected; here the oracle stands in for the GPU context) and the all-gathered result must equal the
single-process result.  Also checks the shard balance rule."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

import kswtest as K


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, seed, out_dir):
    import torch.distributed as dist
    sys.path[:0] = [K.ROOT, os.path.join(K.ROOT, "tests")]
    from bwa_mem_quickassist_b200.sharding import extend_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = K.gen_fuzz(1500, seed=seed, max_q=300)

    def compute(cfg, jobs, qpool, tpool):
        return K.run_oracle(K.Batch(cfg, np.ascontiguousarray(jobs), qpool, tpool), threads=1)

    res = extend_sharded(compute, b.cfg, b.jobs, b.qpool, b.tpool)
    np.save(os.path.join(out_dir, f"res{rank}.npy"), res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_equals_single(tmp_path, oracle_built):
    world, seed = 2, 808
    mp.spawn(_worker, args=(world, _free_port(), seed, str(tmp_path)), nprocs=world, join=True)
    b = K.gen_fuzz(1500, seed=seed, max_q=300)
    want = K.run_oracle(b)
    for r in range(world):
        got = np.load(str(tmp_path / f"res{r}.npy"))
        assert K.first_mismatch(want, got.view(K.RES_DT)) is None


def test_shard_ranges_balance_by_cells():
    from bwa_mem_quickassist_b200.sharding import shard_ranges
    rng = np.random.default_rng(1)
    ql = rng.integers(1, 250, 10000); tl = rng.integers(0, 400, 10000)
    for world in (1, 2, 4, 8):
        rs = shard_ranges(ql, tl, world)
        assert rs[0][0] == 0 and rs[-1][1] == 10000 and all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
        cost = ql.astype(np.int64) * np.maximum(tl, 1)
        sums = np.array([cost[b:e].sum() for b, e in rs])
        assert sums.max() - sums.min() <= 2 * cost.max()
    assert shard_ranges(ql[:1], tl[:1], 4)[-1][1] == 1
