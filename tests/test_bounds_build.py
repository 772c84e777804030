"""The -DIMM3_BOUNDS=1 build of the library (device-side index checks in the block pipeline, k_ptx.cuh: IMM3_CHECK): built by
__graft_entry__.build() next to the product library, loaded here in a child process (IMM3_LIB) and driven through queries that
reach the lane filter kernel, the scan-emit kernel, the general block emit kernel and the offset scan.  compute-sanitizer is
not available on this pool; a failed check traps the kernel and the child fails."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BOUNDS_LIB = os.path.join(ROOT, "immutable3_b200", "libimm3gpu_bounds.so")

CHILD = textwrap.dedent("""
    import sys
    import numpy as np
    sys.path.insert(0, %(root)r)
    sys.path.insert(0, %(root)r + "/tests")
    from helpers import conj
    from immutable3_b200 import Engine, GT, LT, Project, Query, SegmentManager, Select
    from immutable3_b200.loader import SegmentWriter
    d = sys.argv[1]
    rng = np.random.default_rng(1)
    n = 300_000
    ids = (np.arange(n, dtype=np.int64) + (1 << 22) - 5000).astype(np.int32)          # dense sorted: lane kernel fast path, two start widths
    gaps = np.cumsum(rng.integers(0, 9, size=n)).astype(np.int32)                     # mixed widths: quad / per-block routines, emit_any_block
    age = rng.integers(0, 100, size=n).astype(np.int8)
    with SegmentWriter(d, "a", ["id:PFOR_INT", "age:DENSE_TINYINT"], 1024, 40) as w:
        w.append(ids, age)
    with SegmentWriter(d, "b", ["id:PFOR_INT", "age:DENSE_TINYINT"], 96, 500) as w:
        w.append(gaps, age)
    with SegmentManager(d) as sm:
        eng = Engine(sm)
        for name, col in (("a", ids), ("b", gaps)):
            lo, hi = int(col[n // 3]), int(col[2 * n // 3])
            m = (col > lo) & (col < hi)
            for proj in (["id"], ["id", "age"]):
                for limit in (0, 7, 50_000):
                    with eng.execute(Query(name, conj(Select("id", GT(lo)), Select("id", LT(hi))), Project(proj, limit))) as r:
                        exp = col[m] if limit == 0 else col[m][:limit]
                        assert np.array_equal(r.column(0), exp), (name, proj, limit)
            with eng.execute(Query(name, Select("age", LT(10)), Project(["id", "age"]))) as r:   # row-space filter + block emit
                assert np.array_equal(r.column(0), col[age < 10])
    print("bounds build ok")
""")


@pytest.mark.gpu
def test_queries_through_the_bounds_checked_build(tmp_path):
    if not os.path.exists(BOUNDS_LIB):
        pytest.skip("libimm3gpu_bounds.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    env = dict(os.environ, IMM3_LIB=BOUNDS_LIB)
    env.pop("IMM3_PATH", None)
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}, str(tmp_path)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "bounds build ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
