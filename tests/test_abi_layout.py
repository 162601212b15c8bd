"""The ctypes mirror of the C ABI must have exactly the layout the C compiler gives include/otters_b200.h
(sizes and field offsets of every struct that crosses the boundary)."""
import ctypes as C
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STRUCTS = {
    "otters_scan_tuning": "ScanTuning",
    "otters_last_work": "LastWork",
    "otters_vec_query": "VecQuery",
    "otters_column": "Column",
    "otters_build_params": "BuildParams",
    "otters_build_stats": "BuildStats",
    "otters_query_stats": "QueryStats",
    "otters_leaf": "Leaf",
    "otters_filter": "Filter",
    "otters_topk_record": "TopkRecord",
    "otters_shard_map": "ShardMap",
    "otters_peer_exchange": "PeerExchange",
}


def test_ctypes_structs_match_the_header():
    from otters_b200 import _ffi

    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "otters_b200.h"', "int main(void) {"]
    for cname, pyname in STRUCTS.items():
        cls = getattr(_ffi, pyname)
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
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
        cname, size, offs = parts[0], int(parts[1]), [int(x) for x in parts[2:]]
        cls = getattr(_ffi, STRUCTS[cname])
        assert C.sizeof(cls) == size, f"{cname}: ctypes size {C.sizeof(cls)} != C size {size}"
        got = [getattr(cls, fname).offset for fname, _ in cls._fields_]
        assert got == offs, f"{cname}: field offsets {got} != {offs}"
