"""rust/otters-sys/src/lib.rs must declare exactly the ABI of include/otters_b200.h.  No Rust toolchain exists in the authoring
image, so the check is textual but complete: (1) the file is what scripts/gen_rust_sys.py generates from the header today;
(2) independently of that generator, every #[repr(C)] struct is re-parsed from the Rust text, its layout is computed with
the repr(C) rules and compared with gcc's layout of the header (sizes and every field offset, field names in order);
(3) the extern block declares every OTTERS_API symbol with the header's argument count; (4) the shim crate only calls
functions that exist."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RS = os.path.join(ROOT, "rust", "otters-sys", "src", "lib.rs")
SIZES = {"u8": 1, "i8": 1, "u32": 4, "i32": 4, "f32": 4, "c_int": 4, "u64": 8, "i64": 8, "f64": 8, "usize": 8}


def rust_structs():
    text = open(RS).read()
    out = {}
    for m in re.finditer(r"#\[repr\(C\)\]\s*(?:#\[derive\([^)]*\)\]\s*)?pub struct (\w+) \{\n(.*?)\n\}", text, flags=re.S):
        name, body = m.group(1), m.group(2)
        fields = re.findall(r"pub (\w+): ([^,\n]+),", body)
        if fields:
            out[name] = fields
    return out


def repr_c_layout(fields):
    off, align_max, offs = 0, 1, []
    for _, ty in fields:
        ty = ty.strip()
        size = 8 if ty.startswith("*") else SIZES[ty]
        off = (off + size - 1) // size * size
        offs.append(off)
        off += size
        align_max = max(align_max, size)
    return (off + align_max - 1) // align_max * align_max, offs


def test_generated_file_is_current():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_rust_sys.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_rust_struct_layouts_match_gcc():
    structs = rust_structs()
    assert len(structs) >= 12
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "otters_b200.h"', "int main(void) {"]
    for name, fields in structs.items():
        lines.append(f'  printf("{name} %zu", sizeof({name}));')
        for fname, _ in fields:
            lines.append(f'  printf(" %zu", offsetof({name}, {fname}));')  # fails to compile if a field name or order is off
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "layout.c"), os.path.join(d, "layout")
        with open(src, "w") as f:
            f.write("\n".join(lines))
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        parts = line.split()
        name, size, offs = parts[0], int(parts[1]), [int(x) for x in parts[2:]]
        rsize, roffs = repr_c_layout(structs[name])
        assert (rsize, roffs) == (size, offs), f"{name}: Rust repr(C) layout {rsize} {roffs} != C layout {size} {offs}"


def test_struct_field_counts_match_header():
    """Nothing missing at the END of a struct either (offsetof alone would not notice a dropped last field)."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import gen_rust_sys as g

    _, _, c_structs, c_funcs, _, _ = g.parse_header(open(g.HEADER).read())
    rs = rust_structs()
    for name, fields in c_structs:
        assert [f for _, f in fields] == [f for f, _ in rs[name]], name
    text = open(RS).read()
    block = text[text.index('extern "C" {'):]
    rust_fns = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (\w+)\((.*?)\)(?: -> [^;]+)?;", block)}
    assert sorted(rust_fns) == sorted(n for n, _, _ in c_funcs)
    for name, _, args in c_funcs:
        n_rust = 0 if not rust_fns[name].strip() else rust_fns[name].count(":")
        assert n_rust == len(args), name


def test_shim_crate_calls_only_declared_functions():
    text = open(RS).read()
    declared = set(re.findall(r"pub fn (\w+)\(", text)) | set(re.findall(r"pub const (\w+):", text)) | set(re.findall(r"pub struct (\w+)", text))
    shim = open(os.path.join(ROOT, "rust", "otters-gpu", "src", "lib.rs")).read()
    used = set(re.findall(r"sys::(\w+)", shim))
    assert used, "the shim must go through otters-sys"
    assert used <= declared, sorted(used - declared)
