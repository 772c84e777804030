"""The C-ABI shared library loads and exports every symbol include/imm3.h declares (no compute calls)."""
import ctypes
import os
import re

from immutable3_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "imm3.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(imm3_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared_symbols()
    assert len(names) >= 35
    handle = ctypes.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/imm3.h but not exported"
    assert sorted(L.SIGNATURES) == names, "ctypes table and header disagree"
    assert L.lib().imm3_abi_version() == 1


def test_no_oracle_or_cpu_path_in_the_product():
    # the product library must not link or reference the oracle
    blob = open(L.LIB_PATH, "rb").read()
    assert b"orc_query" not in blob and b"liborc" not in blob
    for dirpath, _, files in os.walk(os.path.join(ROOT, "immutable3_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_lib" not in src and "liborc" not in src and "oracle.h" not in src, os.path.join(dirpath, f)


def test_open_without_a_gpu_fails_loudly_not_silently(tmp_path):
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from immutable3_b200 import Imm3Error, SegmentManager

    (tmp_path / "x").mkdir()
    open(tmp_path / "x" / "_table.meta", "w").write('{"name":"x","columns":[{"name":"a","columnType":"INT","codec":"DENSE_INT","dtypeAttrs":{}}],"blockSize":4}')
    with pytest.raises(Imm3Error) as e:
        SegmentManager(tmp_path)
    assert e.value.status == L.ERR_CUDA and "no CPU fallback" in e.value.message


def test_struct_layouts_match_the_jvm_binding():
    """The offsets jvm/Imm3.scala and INTEGRATION.md hard-code (Panama structLayout / p.set(..., offset, ...)) are the C ABI's."""
    P, A, O = L.Pred, L.Agg, L.OpenOpts
    assert (P.col.offset, P.op.offset, P.num.offset, P.strs.offset, P.nstrs.offset, ctypes.sizeof(P)) == (0, 8, 16, 24, 32, 40)
    assert (A.col.offset, A.op.offset, ctypes.sizeof(A)) == (0, 8, 16)
    assert (O.device.offset, O.rank.offset, O.world.offset, O.flags.offset, ctypes.sizeof(O)) == (0, 4, 8, 12, 16)
    scala = open(os.path.join(ROOT, "jvm", "Imm3.scala")).read()
    for sym in ("imm3_open", "imm3_close", "imm3_query", "imm3_query_agg", "imm3_result_nrows", "imm3_result_col_data", "imm3_result_free", "imm3_last_error"):
        assert f'"{sym}"' in scala and sym in L.SIGNATURES, sym
