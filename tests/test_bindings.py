"""The Rust binding ships unbuilt (no cargo in the image): check its extern block against the header and the library."""
import ctypes
import os
import re

from conftest import ROOT


def _c_decls():
    h = open(os.path.join(ROOT, "include", "lrge_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(lrge_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", h, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_rust_extern_block_matches_header_and_library():
    src = open(os.path.join(ROOT, "bindings", "rust", "src", "b200.rs")).read()
    block = re.search(r'extern "C" \{(.*?)\n\}', src, flags=re.S).group(1)
    fns = re.findall(r"fn (lrge_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", block, flags=re.S)
    assert len(fns) >= 6
    decls = _c_decls()
    lib = ctypes.CDLL(os.path.join(ROOT, "lrge_b200", "csrc", "liblrge_b200.so"))
    for name, args in fns:
        assert name in decls, name
        n_rust = len([a for a in args.split(",") if a.strip()])
        assert n_rust == decls[name], (name, n_rust, decls[name])
        assert hasattr(lib, name)
    # struct layouts the binding mirrors
    assert "reserved: [i32; 6]" in src and "int32_t reserved[6];" in open(os.path.join(ROOT, "include", "lrge_b200.h")).read()
