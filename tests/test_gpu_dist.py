"""ShardedEngine with the product CudaExecutor under a real process group (gloo rendezvous, both ranks
on cuda:0 — the data path has no collective, so one GPU is enough to exercise it end to end)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, data_dir, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist

    import oracle_lib as O
    from helpers import conj, oracle_preds
    from immutable3_b200 import GT, LT, Engine, NoSelect, Project, Query, SegmentManager, Select
    from immutable3_b200.dist import CudaExecutor, ShardedEngine

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        with SegmentManager(data_dir, device=0, rank=rank, world=world) as sm, O.Oracle(data_dir) as whole:
            eng = ShardedEngine(CudaExecutor(Engine(sm)), collective_device="cpu")
            for sel, proj in [(conj(Select("age", GT(18)), Select("age", LT(30))), ["id", "age"]),
                              (conj(Select("id", GT(3000)), Select("id", LT(9000))), ["id"]), (NoSelect, ["id", "state"])]:
                for limit in (0, 1, 10, 5000):
                    res = eng.execute(Query("t", sel, Project(proj, limit)))
                    exp = whole.query("t", oracle_preds(sel), proj, limit=limit)
                    assert res.total == exp.nrows
                    for c in range(len(proj)):
                        assert np.array_equal(res.columns[c], exp.columns[c][res.offset:res.offset + res.take])
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_cuda_executor_two_ranks(tmp_path):
    from helpers import make_table

    data = tmp_path / "data"
    make_table(data, "t", 40_000, 64, 5, seed=2)
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, str(data), port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(2))
