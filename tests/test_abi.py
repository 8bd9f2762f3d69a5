"""The C-ABI shared library loads and exports every symbol include/plenvdb_b200.h declares (no compute calls)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "plenvdb_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pvdb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from plenvdb_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 35
    missing = [s for s in syms if not hasattr(_lib.lib, s)]
    assert not missing, "declared in the header but not exported: %s" % missing
    assert not _lib.MISSING_SYMBOLS


def test_ctypes_signatures_cover_the_header():
    from plenvdb_b200 import _lib
    syms = set(_header_symbols())
    bound = set(_lib.DECLARED_SYMBOLS)
    assert syms == bound, "header/binding mismatch: only header %s, only binding %s" % (sorted(syms - bound), sorted(bound - syms))


def test_status_and_error_string_without_gpu():
    from plenvdb_b200 import _lib
    assert _lib.lib.pvdb_abi_version() >= 1
    # argument validation happens before any CUDA call
    try:
        _lib.call("pvdb_zero_grad", None, None, 0, None)
    except _lib.PvdbError as e:
        assert "pvdb_zero_grad" in str(e)
    else:
        raise AssertionError("expected PvdbError")
    h = _lib.lib.pvdb_topo_create_dense(0, 4, 4)
    assert not h and "resolution" in _lib.last_error()


def test_struct_sizes_match_the_header():
    """ctypes mirrors of the PODs must have the C layout (compiled probe)."""
    import ctypes
    import subprocess
    import tempfile
    from plenvdb_b200 import _lib
    code = '#include <stdio.h>\n#include "plenvdb_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(pvdb_tree), ' \
           'sizeof(pvdb_train_cfg), sizeof(pvdb_train_bufs), sizeof(pvdb_render_cfg), sizeof(pvdb_render_bufs), ' \
           'sizeof(pvdb_dp_peers), sizeof(pvdb_frame_peers));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "p.c")
        open(src, "w").write(code)
        exe = os.path.join(d, "p")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    mine = [ctypes.sizeof(c) for c in (_lib.pvdb_tree, _lib.pvdb_train_cfg, _lib.pvdb_train_bufs, _lib.pvdb_render_cfg,
                                       _lib.pvdb_render_bufs, _lib.pvdb_dp_peers, _lib.pvdb_frame_peers)]
    assert sizes == mine, (sizes, mine)
