"""The Rust FFI module of the reference-side shim (integration/b200_sys.rs) is generated from include/b200zk.h:
it must be up to date and declare every exported symbol with the same arity as the header."""
import importlib.util
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen():
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "tools", "gen_rust_ffi.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_generated_file_is_up_to_date():
    g = _gen()
    assert open(g.OUT).read() == g.generate(), "run `python tools/gen_rust_ffi.py`"


def test_every_abi_symbol_is_declared_with_the_headers_arity(zk):
    g = _gen()
    decls = {name: params for name, _ret, params in g.parse_header(open(g.HEADER).read())}
    assert set(decls) == set(zk.ABI_SYMBOLS)
    rust = open(g.OUT).read()
    for name, params in decls.items():
        m = re.search(r"pub fn %s\((.*?)\) -> " % name, rust)
        assert m, name
        assert len([a for a in m.group(1).split(",") if a.strip()]) == len(params), name
    # the argument count of the ctypes binding (the calling convention the tests exercise) agrees as well
    for name, params in decls.items():
        fn = getattr(zk.lib(), name)
        if fn.argtypes is not None:
            assert len(fn.argtypes) == len(params), name


def test_pointer_constness_is_carried_over():
    g = _gen()
    assert g.rust_type("const void* const*") == "*const *const c_void"
    assert g.rust_type("void* const*") == "*const *mut c_void"
    assert g.rust_type("b200zk_srs**") == "*mut *mut Srs"
    assert g.rust_type("const b200zk_ctx*") == "*const Ctx"
    assert g.rust_type("const char*") == "*const c_char"
    assert g.rust_type("uint64_t") == "u64"
