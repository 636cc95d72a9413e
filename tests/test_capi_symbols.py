"""CPU-side boundary checks: the C-ABI library loads, exports every symbol include/openintel_gpu.h
declares, and refuses to work without a GPU (no CPU fallback).  No compute call is made."""
import os
import re

import pytest

import openintel_b200 as oi
from openintel_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "openintel_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(oi_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    return oi.load_library()


def test_library_exports_every_declared_symbol(lib):
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libopenintel_gpu.so does not export " + s
    # and the binding declares a prototype for each of them
    assert sorted(lib._oi_sig) == syms


def test_rust_sys_crate_declares_every_symbol():
    """rust/openintel-gpu-sys is source only (no Rust toolchain here); at least keep its extern block in step
    with the header"""
    rs = open(os.path.join(ROOT, "rust", "openintel-gpu-sys", "src", "lib.rs")).read()
    for s in _declared_symbols():
        assert ("fn %s(" % s) in rs, "rust -sys crate lacks " + s


def test_rust_sources_only_name_symbols_the_header_exports():
    """the port / adapter / use case / MCP tool / CLI sources under rust/openintel-gpu are uncompiled here: every
    `sys::oi_*` function, constant and type they name must exist in the header (functions) or in the -sys crate"""
    sys_rs = open(os.path.join(ROOT, "rust", "openintel-gpu-sys", "src", "lib.rs")).read()
    declared = set(_declared_symbols())
    src_dir = os.path.join(ROOT, "rust", "openintel-gpu", "src")
    files = sorted(f for f in os.listdir(src_dir) if f.endswith(".rs"))
    assert {"lib.rs", "ports.rs", "adapter.rs", "index_builder.rs", "application_search.rs", "mcp_search.rs", "cli_search.rs",
            "wiring.rs"} <= set(files)
    used = set()
    for f in files:
        used |= set(re.findall(r"\bsys::((?:oi|OI)_[A-Za-z0-9_]+)", open(os.path.join(src_dir, f)).read()))
    assert "oi_search_hybrid" in used and "oi_index_create" in used and "oi_lexicon_analyze" in used
    for name in sorted(used):
        if name.startswith("oi_") and name not in ("oi_index", "oi_index_desc", "oi_bm25_params", "oi_status"):
            assert name in declared, "rust source calls %s, which include/openintel_gpu.h does not declare" % name
        assert re.search(r"\b%s\b" % re.escape(name), sys_rs), "rust source names sys::%s, missing from the -sys crate" % name
    # every module lib.rs declares exists, and the use case takes the port as `&dyn HybridSearch` (drop-in injection)
    lib_rs = open(os.path.join(src_dir, "lib.rs")).read()
    for mod in re.findall(r"pub mod ([a-z_]+);", lib_rs):
        assert mod + ".rs" in files, mod
    app = open(os.path.join(src_dir, "application_search.rs")).read()
    assert "searcher: &dyn HybridSearch" in app and "pub async fn search(" in app
    assert "pub struct SearchPostsArgs" in open(os.path.join(src_dir, "mcp_search.rs")).read()
    assert "pub struct SearchArgs" in open(os.path.join(src_dir, "cli_search.rs")).read()


def test_version_and_no_cpu_fallback(lib):
    assert "sm_100a" in oi.version()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device path cannot be exercised")
    with pytest.raises(oi.OiError) as e:
        oi.GpuIndex(n_docs=16, dim=64)
    assert e.value.status == 2 and "no CPU fallback" in e.value.message


def test_bad_descriptor_is_rejected_before_touching_cuda(lib):
    for kw in (dict(dim=6), dict(dim=64, max_k=0), dict(dim=64, max_k=5000), dict(dim=64, dtype=9)):
        args = dict(n_docs=16, dim=64)
        args.update(kw)
        with pytest.raises(oi.OiError) as e:
            oi.GpuIndex(**args)
        assert e.value.status == 1


def test_missing_library_is_a_hard_error(monkeypatch, tmp_path):
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "_HERE", str(tmp_path))
    with pytest.raises(OSError):
        capi.load_library()
