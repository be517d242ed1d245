"""The ctypes mirrors in lrge_b200/__init__.py must have the layout `include/lrge_b200.h` declares: a C program compiled
against the header prints sizeof / offsetof of every struct that crosses the boundary (no GPU, no library needed)."""
import ctypes
import os
import re
import subprocess

import lrge_b200
from lrge_b200 import readset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_layout(tmp_path, struct, fields):
    src = tmp_path / f"layout_{struct}.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "lrge_b200.h"', 'int main(void) {',
             f'  printf("sizeof %zu\\n", sizeof({struct}));']
    lines += [f'  printf("{f} %zu\\n", offsetof({struct}, {f}));' for f in fields]
    lines += ['  return 0;', '}']
    src.write_text("\n".join(lines))
    exe = tmp_path / f"layout_{struct}"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    return dict((k, int(v)) for k, v in (l.split() for l in out.strip().splitlines()))


def _check(tmp_path, struct, cls):
    fields = [f for f, _ in cls._fields_]
    lay = _c_layout(tmp_path, struct, fields)
    assert lay["sizeof"] == ctypes.sizeof(cls), struct
    for f in fields:
        assert lay[f] == getattr(cls, f).offset, (struct, f)


def test_opts_stats_kernel_time_layout(tmp_path):
    _check(tmp_path, "lrge_opts", lrge_b200.Opts)
    _check(tmp_path, "lrge_stats", lrge_b200.Stats)
    _check(tmp_path, "lrge_kernel_time", lrge_b200.KernelTime)


def test_readset_layout(tmp_path):
    _check(tmp_path, "lrge_readset", readset.CReadSet)


def test_overlap_record_layout(tmp_path):
    dt = lrge_b200.OVERLAP_DTYPE
    lay = _c_layout(tmp_path, "lrge_overlap", list(dt.names))
    assert lay["sizeof"] == dt.itemsize
    for f in dt.names:
        assert lay[f] == dt.fields[f][1], f


def test_every_stats_field_is_documented_in_the_header():
    hdr = open(os.path.join(ROOT, "include", "lrge_b200.h")).read()
    body = hdr[hdr.index("typedef struct lrge_stats"):hdr.index("} lrge_stats;")]
    declared = set(re.findall(r"\b([a-z_0-9]+)\s*(?:\[\d+\])?\s*[,;]", body))
    for f, _ in lrge_b200.Stats._fields_:
        assert f in declared, f
