"""The Rust shim (integration/rust/dfd-b200-shim) cannot be compiled here (no cargo/rustc), so its `extern "C"` block is
checked against include/dfd_b200.h as text: every bound function exists in the header with the same number of parameters,
the status constants carry the header's values, and the #[repr(C)] structs list the header's fields in the header's order."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FFI = os.path.join(ROOT, "integration", "rust", "dfd-b200-shim", "src", "ffi.rs")
HEADER = os.path.join(ROOT, "include", "dfd_b200.h")


def _strip_c_comments(text):
    return re.sub(r"/\*.*?\*/", " ", text, flags=re.S)


def _header_functions():
    text = _strip_c_comments(open(HEADER).read())
    out = {}
    for m in re.finditer(r"\b(dfd_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return out


def _rust_functions():
    text = re.sub(r"//[^\n]*", "", open(FFI).read())
    block = text[text.index('extern "C"'):]
    out = {}
    for m in re.finditer(r"pub fn (dfd_\w+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", block, flags=re.S):
        args = [a for a in m.group(2).split(",") if a.strip()]
        out[m.group(1)] = len(args)
    return out


def test_every_bound_function_is_declared_with_the_same_arity():
    header, rust = _header_functions(), _rust_functions()
    assert len(rust) >= 10, rust
    for name, arity in rust.items():
        assert name in header, f"{name} is bound in ffi.rs but not declared in dfd_b200.h"
        assert header[name] == arity, f"{name}: {arity} parameters in ffi.rs, {header[name]} in the header"
    for must in ("dfd_repartition_exec_create", "dfd_repartition_exec_push", "dfd_repartition_exec_finish", "dfd_repartition_exec_abort",
                 "dfd_repartition_exec_execute", "dfd_repartition_exec_destroy", "dfd_schema_supported", "dfd_last_error"):
        assert must in rust, must


def test_status_codes_match_the_header_enum():
    header = _strip_c_comments(open(HEADER).read())
    values = {m.group(1): int(m.group(2)) for m in re.finditer(r"\b(DFD_(?:OK|ERR_\w+))\s*=\s*(\d+)", header)}
    assert values["DFD_OK"] == 0 and len(values) >= 8
    rust = {m.group(1): int(m.group(2)) for m in re.finditer(r"pub const (DFD_\w+): c_int = (\d+);", open(FFI).read())}
    assert set(rust) == set(values), (sorted(rust), sorted(values))
    assert rust == values


def _c_struct_fields(name):
    text = _strip_c_comments(open(HEADER).read())
    m = re.search(r"typedef struct\s*\{([^}]*)\}\s*" + name + r"\s*;", text, flags=re.S)
    assert m, name
    fields = []
    for decl in m.group(1).split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ctype, names = decl.rsplit(None, 1)[0], decl
        ctype = re.match(r"((?:unsigned\s+)?\w+)", decl).group(1)
        for n in decl[len(ctype):].split(","):
            fields.append((n.strip(), ctype))
    return fields


def _rust_struct_fields(name):
    text = open(FFI).read()
    m = re.search(r"pub struct " + name + r"\s*\{(.*?)\}", text, flags=re.S)
    assert m, name
    return [(f.group(1), f.group(2)) for f in re.finditer(r"pub (\w+): (\w+),", m.group(1))]


def test_repr_c_structs_follow_the_header_layout():
    width = {"int64_t": "i64", "int32_t": "i32", "uint64_t": "u64", "uint32_t": "u32"}
    for name in ("dfd_exec_options", "dfd_exec_stats"):
        c, r = _c_struct_fields(name), _rust_struct_fields(name)
        assert [n for n, _ in c] == [n for n, _ in r], (name, c, r)
        assert [width[t] for _, t in c] == [t for _, t in r], (name, c, r)
