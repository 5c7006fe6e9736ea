"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/la_cabi.h declares,
the Python binding table matches the header, and -- with no GPU -- compute entry points fail LOUDLY (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "la_cabi.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"LA_API\s+[\w\s\*]+?\b(la_\w+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for must in ["la_gemm_f64", "la_gemm_f32", "la_gemm_f64_host", "la_gemm_f64_dev", "la_lu_factor_f64",
                 "la_lu_factor_f32", "la_lu_solve_f64", "la_lu_det_f64", "la_lu_is_nonsingular_f64", "la_buf_alloc",
                 "la_buf_upload", "la_buf_download", "la_last_error"]:
        assert must in syms
    assert len(syms) >= 40


def test_library_exports_every_declared_symbol():
    from la import _cabi
    L = ctypes.CDLL(_cabi.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, f"declared in la_cabi.h but not exported: {missing}"


def test_binding_table_matches_header():
    from la import _cabi
    assert sorted(_cabi.SIGNATURES) == declared_symbols()


def test_header_cites_reference_lines():
    text = open(HEADER).read()
    for cite in ["src/matrix/mod.rs:957-980", "src/matrix/mmatrix.rs:82-98", "src/decomp/lu.rs:104-168",
                 "src/decomp/lu.rs:174-182", "src/decomp/lu.rs:224-232", "src/decomp/lu.rs:237-278",
                 "src/internalutil.rs:7-13", "src/matrix/mod.rs:416-426"]:
        assert cite in text, cite


def test_no_cpu_fallback_without_device():
    """In the CPU container every compute call must fail with a clear status, never compute on the host."""
    from la import _cabi
    if _cabi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    L = _cabi.lib()
    a = np.ones(4)
    c = np.zeros(4)
    st = L.la_gemm_f64_host(a.ctypes.data, a.ctypes.data, c.ctypes.data, 2, 2, 2)
    assert st == _cabi.LA_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.la_last_error()
    assert np.all(c == 0)
    piv = np.zeros(2, dtype=np.uint64)
    sign = ctypes.c_int(0)
    st = L.la_lu_factor_f64_host(a.ctypes.data, c.ctypes.data, 2, 2, piv.ctypes.data, ctypes.byref(sign))
    assert st == _cabi.LA_ERR_NO_DEVICE
    h = ctypes.c_void_p()
    assert L.la_buf_alloc(64, 0, ctypes.byref(h)) == _cabi.LA_ERR_NO_DEVICE
    import la
    with pytest.raises(la.LaError):
        la.m("1.0, 2.0; 3.0, 4.0") * la.m("1.0, 0.0; 0.0, 1.0")
    # the multi-device entry points too: no device list can be honoured, and nothing is computed on the host
    devs = (ctypes.c_int * 2)(0, 1)
    st = L.la_lu_factor_f64_mg(2, devs, a.ctypes.data, c.ctypes.data, 2, piv.ctypes.data, ctypes.byref(sign))
    assert st == _cabi.LA_ERR_NO_DEVICE and np.all(c == 0)
    ctx = ctypes.c_void_p()
    assert L.la_lu_mg_create(2, devs, 256, ctypes.byref(ctx)) == _cabi.LA_ERR_NO_DEVICE and not ctx.value
    st = L.la_gemm_f64_mg(2, devs, a.ctypes.data, a.ctypes.data, c.ctypes.data, 2, 2, 2)
    assert st == _cabi.LA_ERR_NO_DEVICE and np.all(c == 0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure; nothing under rust-la_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rust-la_b200")):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".rs", ".c")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"\boracle\b", txt) and "la_fill_hash" not in f:
                    for line in txt.splitlines():
                        if re.search(r"\boracle\b", line) and not re.search(r"//|#|\*|\"\"\"", line):
                            bad.append((f, line.strip()))
    assert not bad, bad


def test_rust_ffi_declares_the_same_functions_as_the_header():
    """rust-la_b200/rust/src/ffi.rs is what the crate binds; there is no Rust toolchain here to compile it,
    so at least its `extern "C"` block must name exactly the functions la_cabi.h declares, with the same argument counts."""
    ffi = open(os.path.join(ROOT, "rust-la_b200", "rust", "src", "ffi.rs")).read()
    rust = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (la_\w+)\s*\(([^)]*)\)", ffi)}
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    cdecl = {m.group(1): m.group(2) for m in re.finditer(r"LA_API\s+[\w\s\*]+?\b(la_\w+)\s*\(([^)]*)\)", text)}
    assert sorted(rust) == sorted(cdecl), sorted(set(rust) ^ set(cdecl))

    def argc(args):
        args = args.strip()
        return 0 if args in ("", "void") else args.count(",") + 1

    wrong = [name for name in cdecl if argc(cdecl[name]) != argc(rust[name])]
    assert not wrong, f"argument count differs between la_cabi.h and ffi.rs: {wrong}"


def test_rust_build_script_compiles_the_same_sources_as_the_makefile():
    """build.rs is what a maintainer's `cargo build` runs; the Makefile is what this repository's tests run.  Both must
    compile the same translation units, or the crate would link against missing symbols."""
    mk = open(os.path.join(ROOT, "rust-la_b200", "Makefile")).read()
    m = re.search(r"^SRCS\s*[:+]?=\s*(.*)$", mk, flags=re.M)
    assert m, "SRCS line not found in the Makefile"
    make_srcs = sorted(os.path.splitext(os.path.basename(w))[0] for w in m.group(1).split() if w.endswith(".cu"))
    rs = open(os.path.join(ROOT, "rust-la_b200", "rust", "build.rs")).read()
    block = re.search(r"let srcs = \[(.*?)\];", rs, flags=re.S).group(1)
    rust_srcs = sorted(re.findall(r'"(\w+)"', block))
    assert rust_srcs == make_srcs, (rust_srcs, make_srcs)


def test_rust_sources_are_at_least_balanced():
    """No Rust toolchain exists in the image, so the mirror cannot be compiled here; this catches the cheapest class of slip
    (unbalanced delimiters after an edit) in rust/src/*.rs and build.rs.  Comments, strings and char literals are skipped."""
    rust_dir = os.path.join(ROOT, "rust-la_b200", "rust")
    files = [os.path.join(rust_dir, "build.rs")] + [os.path.join(rust_dir, "src", f)
                                                     for f in sorted(os.listdir(os.path.join(rust_dir, "src")))]
    pairs = {")": "(", "]": "[", "}": "{"}
    for path in files:
        text = open(path).read()
        text = re.sub(r"//[^\n]*", "", text)
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r'"(?:\\.|[^"\\])*"', '""', text)
        text = re.sub(r"'(?:\\.|[^'\\])'", "' '", text)
        stack = []
        for ch in text:
            if ch in "([{":
                stack.append(ch)
            elif ch in pairs:
                assert stack and stack[-1] == pairs[ch], f"{path}: unbalanced '{ch}'"
                stack.pop()
        assert not stack, f"{path}: unclosed {stack[-3:]}"
