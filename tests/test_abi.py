"""CPU: the C-ABI library builds, loads and exports every symbol include/r4d.h declares (no compute calls)."""
import ctypes
import os

import pytest

from rag4dyg_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build_lib()
    return _lib.load()


def test_header_and_prototypes_agree():
    declared = _lib.header_functions()
    assert declared, "no functions parsed from include/r4d.h"
    assert sorted(_lib.PROTOTYPES) == declared


def test_every_declared_symbol_is_exported(lib):
    for name in _lib.header_functions():
        assert hasattr(lib, name), f"libr4d.so does not export {name}"


def test_pure_host_helpers(lib):
    assert lib.r4d_version() >= 100
    assert lib.r4d_bitset_words(1) == 1 and lib.r4d_bitset_words(32) == 1 and lib.r4d_bitset_words(33) == 2
    assert lib.r4d_bitset_words(20000) == 625
    assert lib.r4d_bitset_pitch_words(20000) == 640      # 2560-byte rows
    assert lib.r4d_bitset_pitch_words(1794) == 64 and lib.r4d_bitset_words(1794) == 57   # UCI_13 history universe
    assert lib.r4d_dense_dpad(768) == 768 and lib.r4d_dense_dpad(512) == 512 and lib.r4d_dense_dpad(100) == 128


def test_argument_errors_are_reported_not_crashing(lib):
    # k out of range is rejected before any device work
    rc = lib.r4d_jaccard_topk(None, None, 0, None, None, 0, 8, 32, 99, 0, 0, 0, None, None, None, None, 0, None)
    assert rc == _lib.R4D_E_ARG
    assert b"k=99" in lib.r4d_last_error()
    rc = lib.r4d_jaccard_topk(None, None, 4, None, None, 4, 8, 32, 10, 0, 0, 0, None, None, None, None, 0, None)
    assert rc == _lib.R4D_E_ARG and b"null pointer" in lib.r4d_last_error()


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.R4DError, match="no CPU fallback"):
        _lib.load()


def test_product_code_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "rag4dyg_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert "oracle" not in src.replace("test oracle", ""), f"{fn} mentions the oracle"


def test_documented_options_exist_with_their_defaults(lib):
    """Every knob include/r4d.h documents for r4d_set_option is known to the library and starts at the documented default;
    an unknown key is an argument error, not a silent no-op."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "r4d.h")).read()
    block = hdr[hdr.index("Tuning / measurement knobs"):hdr.index("int r4d_set_option")]
    knobs = re.findall(r'"(\w+)"\s+\[(-?\d+)\]', block)
    assert len(knobs) == 17 and ("postings_relay", "1") in knobs
    for key, default in knobs:
        if os.environ.get("R4D_" + key.upper()) is not None:
            continue                                  # (environment variables may override the initial value)
        prev = lib.r4d_set_option(key.encode(), int(default))
        assert prev == int(default), f"{key}: header says [{default}], library starts at {prev}"
    assert lib.r4d_set_option(b"no_such_knob", 1) == _lib.R4D_E_ARG
    assert b"unknown key" in lib.r4d_last_error()
