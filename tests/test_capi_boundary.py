"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/mashgpu.h declares, mirrors setAlphabetFromString, and refuses to run without a GPU
(no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mashgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mashgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_three_paths():
    syms = declared_symbols()
    for s in ("mashgpu_sketch_batch", "mashgpu_dist_open", "mashgpu_dist_run", "mashgpu_screen_feed", "mashgpu_screen_finish"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    import mash_b200
    lib = mash_b200.load_library()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_set_alphabet_mirrors_reference_rule(oracle):
    # setAlphabetFromString + use64 rule (reference Sketch.cpp:1108-1137), compared with the oracle's restatement
    import mash_b200
    lib = mash_b200.load_library()
    for k in (1, 8, 15, 16, 17, 21, 32):
        for alpha, pc in (("ACGT", False), ("acgt", False), ("ACDEFGHIKLMNPQRSTVWY", False), ("ACGTN", True)):
            p = mash_b200.SketchParams()
            p.kmer_size = k
            p.preserve_case = int(pc)
            n = lib.mashgpu_set_alphabet(C.byref(p), alpha.encode())
            po = oracle.params(k=k, alphabet=alpha, preserve_case=pc)
            assert n == sum(po.alphabet)
            assert bytes(p.alphabet) == bytes(po.alphabet)
            assert p.use64 == po.use64


def test_no_cpu_fallback_without_gpu():
    import mash_b200
    lib = mash_b200.load_library()
    if lib.mashgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(mash_b200.MashGpuError) as e:
        mash_b200.Engine(0)
    assert "no CUDA device" in str(e.value) or "CPU" in str(e.value)


def test_product_never_imports_oracle():
    # the product path must not route through the checker
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mash_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.replace("# oracle-free", ""), os.path.join(dirpath, f)


def test_every_environment_switch_of_the_library_is_documented():
    # INTEGRATION.md section 5 lists the switches; a getenv() added to the sources without a line there fails here
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    used = set()
    for sub in ("mash_b200/csrc", "mash_b200/host"):
        d = os.path.join(root, sub)
        for name in os.listdir(d):
            if name.endswith((".cu", ".cuh", ".cpp", ".hpp", ".h")):
                used.update(re.findall(r'getenv\("(MASH[A-Z_0-9]*)"\)', open(os.path.join(d, name), errors="replace").read()))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    missing = sorted(v for v in used if v not in doc)
    assert used and not missing, missing
