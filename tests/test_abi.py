"""CPU: the C-ABI library loads, exports every symbol include/smow_b200.h declares, and the Python
operators refuse to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from helpers import ROOT
from smow_net_b200 import _lib, ops


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "smow_b200.h")).read()
    return sorted(set(re.findall(r"\b(smow_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the extension: python -m smow_net_b200.build"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_version_and_options_without_gpu():
    lib = _lib.load()
    assert lib.smow_abi_version() == _lib.ABI_VERSION
    old = _lib.get_option("warp_fwd_variant")
    _lib.set_option("warp_fwd_variant", 0)
    assert _lib.get_option("warp_fwd_variant") == 0
    _lib.set_option("warp_fwd_variant", old)
    with pytest.raises(RuntimeError, match="unknown option"):
        _lib.set_option("no_such_knob", 1)
    assert _lib.get_option("no_such_knob") == -1
    assert isinstance(_lib.launch_count(), int)


def test_argument_errors_are_reported_not_thrown():
    lib = _lib.load()
    rc = lib.smow_warp_stack_fwd(None, None, None, None, None, 1, 4, 8, 8, 0, 0, None)
    assert rc == -1 and b"null" in lib.smow_last_error()
    rc = lib.smow_tlerp_cat_fwd(None, None, None, 1, 0, 4, 64, 0, 0, None)
    assert rc == -1


def test_operators_refuse_cpu_tensors():
    x = torch.randn(1, 4, 2, 8, 8)
    flow = torch.zeros(1, 2, 2, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.flow_warp(x, flow, (8, 8))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.tlerp(x)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.tlerp_pair_cat(None, x[:, :, 0], x[:, :, 1])


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libsmow_b200.so")
    with pytest.raises(RuntimeError, match="mandatory"):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "smow_net_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "liboracle" not in src, f


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the reference's own modules from oracle/_ref/reference.zip on the host CPU; the oracle's
    port only where that archive is absent) prints ONE JSON line with the contract's keys; a small batch keeps it short."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--batch", "2"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["value"] > 0
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    from oracle import ref_runtime
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_runtime.available() else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["nproc"] >= 1 and d["cpu_baseline"]["cpu_model"]
    assert len([l for l in r.stdout.splitlines() if l.strip()]) == 1      # stdout is exactly the one JSON line
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_widened_operators_have_no_cpu_path_either():
    """Rows N2 / N4: like the warp and lerp operators, ops.semantic_tokens and ops.frame_mix refuse CPU tensors instead
    of falling back (the modules keep the reference's own PyTorch composition for CPU tensors; the operators do not)."""
    x = torch.zeros(1, 16, 4, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.semantic_tokens(x, torch.zeros(8, 16, 1, 1), torch.zeros(8))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.frame_mix(x, torch.zeros(16, 16), torch.zeros(4, 16, 16))
    assert _lib.load().smow_frame_mix_supported(28) == 1 and _lib.load().smow_frame_mix_supported(24) == 0
    assert _lib.load().smow_tokenizer_workspace_bytes(2, 16, 16384) == 4 * 2 * 32 * (16 + 8 * 16) * 4


def test_fused_warp_tokens_has_no_cpu_path_and_reports_argument_errors():
    """Rows A1 + N2 fused: ops.warp_tokens refuses CPU tensors (the modules then run flow_warp + the tokenizer loop, which the
    CPU tests route to the oracle); the C entry points report bad arguments through their return code; the support query and
    the knob of the row-wise BatchNorm backward work without a GPU."""
    x, flow = torch.zeros(1, 16, 2, 8, 8), torch.zeros(1, 2, 2, 8, 8)
    w, b = torch.zeros(8, 16, 1, 1), torch.zeros(8)
    assert not ops.warp_tokens_supported(x, w)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.warp_tokens(x, flow, w, b)
    lib = _lib.load()
    assert lib.smow_warp_tokenizer_supported(16) == 1 and lib.smow_warp_tokenizer_supported(32) == 1
    assert lib.smow_warp_tokenizer_supported(64) == 0 and lib.smow_warp_tokenizer_supported(28) == 0
    old = _lib.get_option("tok_variant")
    _lib.set_option("tok_variant", 0)                       # the fused pass exists only in the tensor-core family
    try:
        assert lib.smow_warp_tokenizer_supported(16) == 0
    finally:
        _lib.set_option("tok_variant", old)
    rc = lib.smow_warp_tokenizer_fwd(None, None, None, None, None, None, None, None, 1, 16, 8, 8, 0, 1, None, 0, None)
    assert rc == -1 and lib.smow_last_error()
    rc = lib.smow_warp_tokenizer_bwd(None, None, None, None, None, None, None, None, None, None, None, None, 1, 16, 0, 8, 0, 1,
                                     None, 0, None)
    assert rc == -1
    assert _lib.get_option("bn_bwd_rows") == 1
    _lib.set_option("bn_bwd_rows", 0)
    assert _lib.get_option("bn_bwd_rows") == 0
    _lib.set_option("bn_bwd_rows", 1)
