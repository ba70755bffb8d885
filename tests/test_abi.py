"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/sdfb.h
declares, argument validation mirrors the reference, and there is no CPU fallback."""
import ctypes
import os
import re

import numpy as np
import pytest

import sdfgen_b200
from sdfgen_b200 import _lib, meshes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    return sdfgen_b200.is_gpu_available()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sdfb.h")).read()
    names = sorted(set(re.findall(r"\b(sdfb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 15
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), n
    assert b"sm_100a" in _lib.lib().sdfb_version()


def test_library_has_only_sm100a_code():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_validation_errors_match_reference_messages():
    v, t = meshes.unit_cube()
    o = (0.0, 0.0, 0.0)
    with pytest.raises(ValueError, match="empty mesh"):
        sdfgen_b200.generate_sdf(np.zeros((0, 3), np.float32), t, o, 0.1, 8, 8, 8)
    with pytest.raises(ValueError, match="empty mesh"):
        sdfgen_b200.generate_sdf(v, np.zeros((0, 3), np.uint32), o, 0.1, 8, 8, 8)
    with pytest.raises(ValueError, match="must be positive"):
        sdfgen_b200.generate_sdf(v, t, o, 0.1, 0, 8, 8)
    with pytest.raises(ValueError, match="must be positive"):
        sdfgen_b200.generate_sdf(v, t, o, 0.1, 8, -1, 8)
    with pytest.raises(ValueError, match="dx must be positive"):
        sdfgen_b200.generate_sdf(v, t, o, 0.0, 8, 8, 8)
    with pytest.raises(ValueError, match="Invalid backend"):
        sdfgen_b200.generate_sdf(v, t, o, 0.1, 8, 8, 8, backend="tpu")
    with pytest.raises(ValueError, match="no CPU fallback"):
        sdfgen_b200.generate_sdf(v, t, o, 0.1, 8, 8, 8, backend="cpu")


def test_no_cpu_fallback_without_device():
    if _has_gpu():
        pytest.skip("a GPU is present")
    assert _lib.lib().sdfb_device_count() == 0
    v, t = meshes.unit_cube()
    with pytest.raises(_lib.SdfbError) as e:
        sdfgen_b200.generate_sdf(v, t, (0, 0, 0), 0.1, 8, 8, 8)
    assert e.value.code == _lib.ERR_NO_DEVICE
    with pytest.raises(_lib.SdfbError):
        _lib.Plan(8, 8, 8)


def test_c_abi_argument_checks_without_device():
    L = _lib.lib()
    h = ctypes.c_void_p()
    assert L.sdfb_plan_create(ctypes.byref(h), 0, 0, 8, 8, 0, 8, 0) == _lib.ERR_INVALID
    assert b"positive" in L.sdfb_last_error()
    assert L.sdfb_plan_create(ctypes.byref(h), 0, 8, 8, 8, 4, 2, 0) == _lib.ERR_INVALID
    assert L.sdfb_plan_create(None, 0, 8, 8, 8, 0, 8, 0) == _lib.ERR_INVALID
    assert L.sdfb_plan_destroy(None) == 0
    assert L.sdfb_plan_band(None, None, 1.0, 1, None) == _lib.ERR_INVALID
    assert L.sdfb_make_level_set3(None, 0, None, 0, None, 1.0, 4, 4, 4, 1, None, None, None, 0) == _lib.ERR_INVALID
    # entry points added for SURVEY 8(f): writer, batch, concurrency, memory pool
    assert L.sdfb_plan_write_sdf(None, b"/tmp/x.sdf", None, 1.0, None, None) == _lib.ERR_INVALID
    assert L.sdfb_plan_set_concurrency(None, 2) == _lib.ERR_INVALID
    assert L.sdfb_make_level_set3_batch(None, 1, 4, 0) == _lib.ERR_INVALID
    assert L.sdfb_make_level_set3_batch(None, -1, 4, 0) == _lib.ERR_INVALID
    assert L.sdfb_make_level_set3_batch(None, 0, 4, 0) == 0            # an empty batch is not an error
    assert L.sdfb_trim_memory() == 0                                   # nothing retained yet


def test_batch_item_layout_matches_the_header():
    """sdfb_batch_item as ctypes sees it = as the C compiler lays it out (offsets from include/sdfb.h compiled with gcc)."""
    import subprocess, tempfile, os
    fields = ["tri", "ntri", "xyz", "nvert", "origin", "dx", "ni", "nj", "nk", "exact_band", "phi_out", "status"]
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "sdfb.h"\nint main(void){printf("%zu", sizeof(sdfb_batch_item));' + \
          "".join(f'printf(" %zu", offsetof(sdfb_batch_item, {f}));' for f in fields) + "return 0;}\n"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", inc, "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        out = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    assert out[0] == ctypes.sizeof(_lib.BatchItem)
    assert out[1:] == [getattr(_lib.BatchItem, f).offset for f in fields]


def test_ctypes_prototypes_match_the_header():
    """Every prototype in include/sdfb.h against what sdfgen_b200/_lib.py tells ctypes: same number of parameters, and per
    parameter the same class (pointer / 64-bit integer / 32-bit integer / float) -- a drifted binding would pass garbage
    through the C ABI without any error."""
    hdr = open(os.path.join(ROOT, "include", "sdfb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = re.findall(r"\b(int|uint64_t|const char \*)\s*(sdfb_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr)
    assert len(protos) >= 30

    def c_class(param):
        p = " ".join(param.split())
        if p == "void":
            return None
        if "*" in p or "[" in p:
            return "ptr"
        if p.startswith(("uint64_t", "int64_t")):
            return "i64"
        if p.startswith(("int32_t", "uint32_t", "int ")):
            return "i32"
        if p.startswith("float"):
            return "f32"
        raise AssertionError("unclassified parameter: " + p)

    def ct_class(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or issubclass(t, ctypes._Pointer):
            return "ptr"
        return {ctypes.c_uint64: "i64", ctypes.c_int64: "i64", ctypes.c_int32: "i32", ctypes.c_uint32: "i32",
                ctypes.c_int: "i32", ctypes.c_float: "f32"}[t]

    L = _lib.lib()
    checked = 0
    for ret, name, params in protos:
        want = [c for c in (c_class(x) for x in params.split(",")) if c]
        f = getattr(L, name)
        if f.argtypes is None:
            assert not want or name in ("sdfb_version", "sdfb_last_error", "sdfb_device_count", "sdfb_launch_count"), name
            continue
        assert [ct_class(t) for t in f.argtypes] == want, (name, want, f.argtypes)
        want_ret = {"int": ctypes.c_int, "uint64_t": ctypes.c_uint64, "const char *": ctypes.c_char_p}[ret]
        assert f.restype is want_ret, (name, f.restype)
        checked += 1
    assert checked >= 28


def test_missing_library_fails_loudly():
    """Without libsdfb.so the package must not compute anything: the first call is an ImportError that says how to build
    the extension, from every public entry (there is no Python, torch or oracle fallback behind it)."""
    import subprocess
    import sys
    code = (
        "import numpy as np, sdfgen_b200\n"
        "from sdfgen_b200 import meshes\n"
        "v, t = meshes.unit_cube()\n"
        "calls = [lambda: sdfgen_b200.is_gpu_available(), lambda: sdfgen_b200.generate_sdf(v, t, (0, 0, 0), 0.1, 8, 8, 8),\n"
        "         lambda: sdfgen_b200.generate_from_mesh(v, t, nx=8), lambda: sdfgen_b200.Plan(8, 8, 8),\n"
        "         lambda: sdfgen_b200.generate_sdf_batch([dict(vertices=v, triangles=t, origin=(0, 0, 0), dx=0.1, nx=4, ny=4, nz=4)])]\n"
        "for c in calls:\n"
        "    try:\n"
        "        c()\n"
        "    except ImportError as e:\n"
        "        assert 'no CPU fallback' in str(e) and 'make -C sdfgen_b200/csrc' in str(e)\n"
        "    else:\n"
        "        raise SystemExit('a call succeeded without the library')\n"
        "import sys\n"
        "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'the product imported the oracle'\n"
        "print('LOUD')\n")
    env = dict(os.environ, SDFB_LIB_PATH="/nonexistent/libsdfb.so", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and "LOUD" in r.stdout, r.stdout + r.stderr


def test_product_sources_never_reach_into_the_oracle():
    """Static side of the same rule: no module under sdfgen_b200/ (or its alias package) imports `oracle`, the library's
    build does not compile or link anything from oracle/, and libsdfb.so does not depend on the checkers' shared objects."""
    import ast
    import subprocess
    for pkg in ("sdfgen_b200", "sdfgenfast_b200"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                path = os.path.join(dirpath, f)
                if f.endswith(".py"):
                    for node in ast.walk(ast.parse(open(path).read())):
                        names = [a.name for a in node.names] if isinstance(node, ast.Import) else \
                                [node.module or ""] if isinstance(node, ast.ImportFrom) else []
                        assert not any(n == "oracle" or n.startswith("oracle.") for n in names), path
                elif f.endswith((".cu", ".cuh", ".h", ".hpp")) or f == "Makefile":
                    text = open(path).read()
                    assert not re.search(r'#include\s*[<"][^>"]*oracle', text), path
                    assert "liboracle" not in text and "libsdfgen_ref" not in text, path
    r = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True)
    if r.returncode == 0:
        assert "oracle" not in r.stdout and "sdfgen_ref" not in r.stdout and "libtorch" not in r.stdout and "libcudart" not in r.stdout, r.stdout
