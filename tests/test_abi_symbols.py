"""CPU: the C-ABI shared library loads and exports every symbol include/hgru_b200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hgru_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"HGRU_API\s+[\w\s\*]+?\b(\w+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = _declared()
    for must in ("hgru_plan_create", "hgru_set_params", "hgru_forward", "hgru_plan_destroy",
                 "pose_plan_create", "pose_set_params", "pose_forward", "pose_forward_host",
                 "pose_plan_destroy", "hgru_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from monkey_pose_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        pytest.fail("libhgru_b200.so missing: run __graft_entry__.build()")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert set(_declared()) == set(_lib.SIGNATURES), "ctypes table out of sync with the header"


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md says, for every exported symbol, which reference interface it replaces (or that it is
    introspection only)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in _declared() if n not in doc]
    assert not missing, missing


def test_typed_load_and_pure_host_calls():
    from monkey_pose_b200 import _lib
    lib = _lib.load()
    assert lib.hgru_version() >= 100
    assert lib.hgru_plan_destroy(None) == 0          # destroying NULL is a no-op
    assert lib.hgru_plan_workspace_bytes(None) == 0
    # argument validation happens before any CUDA call
    assert lib.hgru_plan_create(1, 8, 8, 8, 15, 1, 0, None) == 1
    assert b"out is null" in lib.hgru_last_error()
    h = ctypes.c_void_p()
    assert lib.hgru_plan_create(1, 8, 8, 8, 4, 1, 0, ctypes.byref(h)) == 2     # even filter size
    assert lib.hgru_plan_create(0, 8, 8, 8, 15, 1, 0, ctypes.byref(h)) == 1    # empty batch
    assert lib.hgru_plan_create(1, 8, 8, 8, 15, 1, 7, ctypes.byref(h)) == 2    # unknown mode


def test_pose_params_struct_matches_header_layout():
    from monkey_pose_b200 import _lib
    # 10 stem/fc pointers + 5x4 batch-norm pointers + 12 hGRU pointers
    assert ctypes.sizeof(_lib.PoseParams) == (10 + 20 + 12) * ctypes.sizeof(ctypes.c_void_p)


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/hgru_b200.h must compile as C99 without warnings and a C program
    must link against the library (no C++ or torch types in the signatures)."""
    import os
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "monkey-pose_b200", "libhgru_b200.so")
    if not os.path.exists(lib):
        import pytest
        pytest.skip("library not built")
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "abi.c")
        open(src, "w").write('#include "hgru_b200.h"\n#include <stdio.h>\n'
                             'int main(void) { printf("%d\\n", hgru_version()); return 0; }\n')
        exe = os.path.join(d, "abi")
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                            src, "-o", exe, lib, "-Wl,-rpath," + os.path.dirname(lib)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0 and int(out.stdout.strip()) >= 100, (out.stdout, out.stderr)


def test_documents_cite_only_evidence_files_that_exist():
    """Every `r01_… / r02_…` evidence file named in DESIGN.md, README.md, INTEGRATION.md and profiles/README.md is
    present under profiles/ (shell-style wildcards allowed)."""
    import glob
    import re
    missing = []
    for doc in ("DESIGN.md", "README.md", "INTEGRATION.md", os.path.join("profiles", "README.md")):
        text = open(os.path.join(ROOT, doc)).read()
        for m in re.finditer(r"`((?:profiles/)?r0[12]_[A-Za-z0-9_.*?-]+)`", text):
            name = m.group(1)
            path = os.path.join(ROOT, name if name.startswith("profiles/") else os.path.join("profiles", name))
            if not glob.glob(path):
                missing.append((doc, name))
    assert not missing, missing
