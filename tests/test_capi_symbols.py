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


def test_ctypes_argument_counts_match_header_prototypes():
    """Every ctypes binding in _lib.py takes exactly as many arguments as the header's prototype declares (a mismatch
    would only show up as garbage arguments on the GPU)."""
    from d3roma_b200 import _lib
    l = _lib.lib()
    text = (ROOT / "include" / "d3roma_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"\s+", " ", text)
    protos = re.findall(r"\b(d3_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", text)
    assert len(protos) >= 90
    checked = 0
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        fn = getattr(l, name)
        if fn.argtypes is None:
            assert n == 0, f"{name}: {n} parameters in the header but no argtypes bound"
            continue
        assert len(fn.argtypes) == n, f"{name}: header declares {n} parameters, ctypes binds {len(fn.argtypes)}"
        checked += 1
    assert checked >= 85


def test_ctypes_argument_types_match_header_prototypes():
    """Scalar classes of every bound argument (pointer / float / double / 32-bit / 64-bit / size_t) against the header."""
    import ctypes as C
    from d3roma_b200 import _lib
    l = _lib.lib()
    text = (ROOT / "include" / "d3roma_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"\s+", " ", text)

    def header_class(param: str) -> str:
        if "*" in param or "[" in param:
            return "ptr"
        for tok, cls in (("double", "f64"), ("float", "f32"), ("int64_t", "i64"), ("size_t", "size"), ("uint32_t", "u32"),
                         ("int32_t", "i32"), ("int", "i32")):
            if re.search(r"\b%s\b" % tok, param):
                return cls
        raise AssertionError(f"unclassified parameter {param!r}")

    def ctypes_class(t) -> str:
        if t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or issubclass(t, C._Pointer) or issubclass(t, C.Array):
            return "ptr"
        return {C.c_float: "f32", C.c_double: "f64", C.c_int64: "i64", C.c_size_t: "size", C.c_uint32: "u32",
                C.c_int32: "i32", C.c_int: "i32"}[t]

    same_width = {("size", "i64"), ("i64", "size")}            # LP64: both are 64-bit integers in the same register class
    bad = []
    for name, params in re.findall(r"\b(d3_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", text):
        fn = getattr(l, name)
        if fn.argtypes is None:
            continue
        want = [] if params.strip() in ("", "void") else [header_class(p.strip()) for p in params.split(",")]
        got = [ctypes_class(t) for t in fn.argtypes]
        for i, (w, g) in enumerate(zip(want, got)):
            if w != g and (w, g) not in same_width:
                bad.append(f"{name} arg {i}: header {w}, ctypes {g}")
    assert not bad, bad


def test_every_exported_entry_point_is_declared_in_the_header():
    """The reverse direction: no ``d3_*`` function leaves the library without a prototype (and its reference citation)
    in include/d3roma_b200.h."""
    import shutil
    import subprocess
    from d3roma_b200 import _lib
    nm = shutil.which("nm")
    if nm is None:
        import pytest
        pytest.skip("nm not available")
    out = subprocess.run([nm, "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = sorted({line.split()[-1] for line in out.splitlines() if line.split() and line.split()[-1].startswith("d3_")
                       and line.split()[-2] in ("T", "t")})
    assert len(exported) >= 90
    undeclared = [s for s in exported if s not in set(_declared())]
    assert not undeclared, undeclared
