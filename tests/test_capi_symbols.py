"""The C-ABI shared library must load without a GPU and export every symbol include/d3roma_b200.h declares
(no compute calls here)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "d3roma_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(d3_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_all_declared_symbols():
    from d3roma_b200 import _lib, build
    if not _lib.LIB_PATH.exists():
        build.build(verbose=False)
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    names = _declared()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.d3_version.restype = ctypes.c_int
    assert lib.d3_version() >= 100


def test_ctypes_signatures_cover_header():
    from d3roma_b200 import _lib
    l = _lib.lib()  # _declare() raises AttributeError if a bound symbol is missing from the library
    for n in _declared():
        assert getattr(l, n).argtypes is not None or n in ("d3_last_error", "d3_version", "d3_launch_count",
                                                            "d3_reset_launch_count"), n


def test_struct_layouts_match_header():
    """Field order of the ctypes mirrors == field order in the header structs."""
    from d3roma_b200 import _lib
    text = (ROOT / "include" / "d3roma_b200.h").read_text()
    for cname, cls in (("d3_conv_params", _lib.ConvParams), ("d3_wgrad_params", _lib.WgradParams),
                       ("d3_pack_entry", _lib.PackEntry),
                       ("d3_ddim_coeffs", _lib.DDIMCoeffs),
                       ("d3_ddpm_coeffs", _lib.DDPMCoeffs), ("d3_guidance_params", _lib.GuidanceParams)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), text, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(void|float|double|int32_t|int64_t)\s*\*?", "", decl).strip()
            fields += [re.sub(r"\[\d+\]$", "", f.strip().lstrip("*").strip()) for f in decl.split(",")]
        assert fields == [f[0] for f in cls._fields_], (cname, fields)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from d3roma_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    import pytest
    with pytest.raises(_lib.D3Error):
        _lib.lib()
