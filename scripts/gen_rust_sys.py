#!/usr/bin/env python3
"""Generates rust/otters-sys/src/lib.rs from include/otters_b200.h (no bindgen and no Rust toolchain in the authoring image:
the ABI is small, so a 150-line translator of the header's `#define`s, `typedef struct`s and `OTTERS_API` prototypes does).

    python scripts/gen_rust_sys.py            # rewrite rust/otters-sys/src/lib.rs
    python scripts/gen_rust_sys.py --check    # exit 1 if the committed file is stale

tests/test_rust_sys.py runs the check and, independently of this translator, recomputes every struct's repr(C) layout from
the Rust text and compares it with the ctypes structures (which tests/test_abi_layout.py pins to gcc's offsets)."""
from __future__ import annotations

import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "otters_b200.h")
OUT = os.path.join(ROOT, "rust", "otters-sys", "src", "lib.rs")

SCALARS = {
    "void": "c_void", "char": "c_char", "int": "c_int", "float": "f32", "double": "f64", "uint8_t": "u8", "uint32_t": "u32",
    "int32_t": "i32", "uint64_t": "u64", "int64_t": "i64", "size_t": "usize",
}


def strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def rust_type(decl: str, structs) -> str:
    """C type (without the declared name) -> Rust type."""
    toks = re.findall(r"\*|\w+", decl)
    base, base_const, levels = None, False, []
    for t in toks:
        if t == "const":
            if base is None or not levels:
                if base is None:
                    base_const = True
                else:
                    base_const = True  # `T const`
            else:
                levels[-1] = True
        elif t == "*":
            levels.append(False)
        elif t in ("struct",):
            continue
        else:
            base = t
    assert base is not None, decl
    out = SCALARS.get(base, base if base in structs else None)
    assert out is not None, f"unknown C type {base!r} in {decl!r}"
    for i in range(len(levels)):
        pointee_const = base_const if i == 0 else levels[i - 1]
        out = ("*const " if pointee_const else "*mut ") + out
    return out


def split_decl(field: str):
    """'const float *queries' -> ('const float *', 'queries')"""
    m = re.match(r"^(.*?)(\w+)\s*$", field.strip(), flags=re.S)
    return m.group(1).strip(), m.group(2)


def parse_header(text: str):
    raw = text
    text = strip_comments(text)
    defines = [(m.group(1), m.group(2)) for m in re.finditer(r"^#define\s+(OTTERS_[A-Z0-9_]+)\s+(-?\d+)\s*$", text, flags=re.M)]
    opaque = re.findall(r"typedef\s+struct\s+(\w+)\s+\1\s*;", text)
    structs = []
    for m in re.finditer(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for f in m.group(1).split(";"):
            f = " ".join(f.split())
            if not f:
                continue
            parts = [x.strip() for x in f.split(",")]  # `float a, b, c;` declares three fields of one type
            t0, n0 = split_decl(parts[0])
            fields.append((t0, n0))
            for extra in parts[1:]:
                assert "*" not in extra, f
                fields.append((t0, extra))
        structs.append((m.group(2), fields))
    names = set(opaque) | {n for n, _ in structs}
    funcs = []
    for m in re.finditer(r"OTTERS_API\s+(.*?)\(([^;]*?)\)\s*;", text, flags=re.S):
        head = " ".join(m.group(1).split())
        ret, name = split_decl(head)
        args = []
        arg_text = " ".join(m.group(2).split())
        if arg_text and arg_text != "void":
            for a in arg_text.split(","):
                args.append(split_decl(a))
        funcs.append((name, ret, args))
    return defines, opaque, structs, funcs, names, raw


def generate() -> str:
    defines, opaque, structs, funcs, names, _ = parse_header(open(HEADER).read())
    o = []
    o.append("//! Raw FFI declarations of libotters_b200.so — GENERATED from include/otters_b200.h by scripts/gen_rust_sys.py; do not edit.")
    o.append("//! Every item mirrors the C header one to one; see the header for the reference lines each entry point replaces.")
    o.append("//! tests/test_rust_sys.py keeps this file in step with the header and checks every repr(C) layout.")
    o.append("#![allow(non_camel_case_types)]")
    o.append("use std::os::raw::{c_char, c_int, c_void};")
    o.append("")
    for n, v in defines:
        o.append(f"pub const {n}: c_int = {v};")
    o.append("")
    for n in opaque:
        o.append("#[repr(C)]")
        o.append(f"pub struct {n} {{ _private: [u8; 0] }}")
    o.append("")
    for n, fields in structs:
        all_scalar = all("*" not in t for t, _ in fields)
        o.append("#[repr(C)]")
        o.append("#[derive(Clone, Copy, Default)]" if all_scalar else "#[derive(Clone, Copy)]")
        o.append(f"pub struct {n} {{")
        for t, f in fields:
            o.append(f"    pub {f}: {rust_type(t, names)},")
        o.append("}")
        o.append("")
    o.append('extern "C" {')
    for name, ret, args in funcs:
        a = ", ".join(f"{('r#' + an) if an in ('type', 'match', 'ref') else an}: {rust_type(at, names)}" for at, an in args)
        r = rust_type(ret, names)
        o.append(f"    pub fn {name}({a})" + ("" if r == "c_void" else f" -> {r}") + ";")
    o.append("}")
    return "\n".join(o) + "\n"


def main():
    text = generate()
    if "--check" in sys.argv:
        cur = open(OUT).read() if os.path.exists(OUT) else ""
        if cur != text:
            print("rust/otters-sys/src/lib.rs is stale: run python scripts/gen_rust_sys.py", file=sys.stderr)
            sys.exit(1)
        return
    with open(OUT, "w") as f:
        f.write(text)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
