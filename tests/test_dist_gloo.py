"""The N>1 path on CPU: world_size-2 (and 3) gloo process groups drive ShardedEngine — shard slicing,
the count all-gather, the offset/LIMIT split and the ordered concatenation.  The local executor is
injected: here it is backed by the ORACLE (tests may use it as a stand-in), because there is no GPU in
this container and the product has no CPU path; on the GPU box the same class runs with CudaExecutor
(tests/test_gpu_dist.py)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


class OracleExecutor:
    """begin/finish over the oracle restricted to this rank's canonical slice."""

    def __init__(self, data_dir, rank, world):
        import oracle_lib as O
        from immutable3_b200.dist import shard_range

        self.orc = O.Oracle(data_dir)
        self.rank, self.world = rank, world
        self.shard_range = shard_range

    def begin(self, query):
        from helpers import oracle_preds

        a, b = self.shard_range(self.orc.nsegments(query.table), self.rank, self.world)
        r = self.orc.query(query.table, oracle_preds(query.select), list(query.project.cols), limit=query.project.limit, seg_begin=a, seg_end=b)
        r.local_count = r.nrows
        return r

    def finish(self, handle, take):
        return [c[:take] for c in handle.columns]


def _worker(rank, world, data_dir, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist

    import oracle_lib as O
    from helpers import conj, oracle_preds
    from immutable3_b200 import GT, LT, Match, NoSelect, Project, Query, Select
    from immutable3_b200.dist import ShardedEngine

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        eng = ShardedEngine(OracleExecutor(data_dir, rank, world))
        whole = O.Oracle(data_dir)
        cases = [("t", conj(Select("age", GT(18)), Select("age", LT(30))), ["id", "age"]),
                 ("t", conj(Select("id", GT(300)), Select("id", LT(900))), ["id"]),          # matches live on few ranks: skewed output
                 ("t", Select("state", Match(["CA"])), ["state", "id"]),
                 ("t", NoSelect, ["id"])]
        for table, sel, proj in cases:
            for limit in (0, 1, 7, 50, 10**6):
                res = eng.execute(Query(table, sel, Project(proj, limit)))
                exp = whole.query(table, oracle_preds(sel), proj, limit=limit)
                assert res.total == exp.nrows, (table, limit, res.counts)
                assert res.take == len(res.columns[0])
                # every rank can check its own slice of the global answer without any row exchange
                for c in range(len(proj)):
                    assert np.array_equal(res.columns[c], exp.columns[c][res.offset:res.offset + res.take]), (rank, table, limit)
                rows = eng.gather_rows(res)
                if rank == 0:
                    for c in range(len(proj)):
                        assert np.array_equal(rows[c], exp.columns[c])
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_engine_over_gloo(tmp_path, world):
    from helpers import make_table

    data = tmp_path / "data"
    make_table(data, "t", 13 * (8 * 2 + 1) + 5, 8, 2, seed=2)  # 14 segments: lexicographic canonical order
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, str(data), port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
