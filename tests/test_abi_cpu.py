"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU or NCCL and exports every
symbol include/tnet_b200.h declares; compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import subprocess

import pytest

from tnet_b200 import abi


def test_library_is_built():
    assert os.path.exists(abi.LIB_PATH), "run `python -c 'import __graft_entry__ as g; g.build()'`"


def test_exports_every_declared_symbol():
    lib = abi.lib()
    names = abi.declared_symbols()
    assert len(names) >= 55
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_link_time_dependency_on_driver_or_nccl_or_oracle():
    out = subprocess.check_output(["ldd", abi.LIB_PATH], text=True)
    for forbidden in ("libcuda.so", "libnccl", "libcublas", "tnet_oracle"):
        assert forbidden not in out, out


def test_fails_loudly_without_gpu():
    lib = abi.lib()
    n = C.c_int(-1)
    rc = lib.tnb_device_count(C.byref(n))
    if rc == abi.OK and n.value > 0:
        pytest.skip("a GPU is visible")
    h = C.c_void_p()
    rc = lib.tnb_ctx_create(C.byref(h), C.c_int(0))
    assert rc != abi.OK and not h.value
    assert lib.tnb_last_error().decode() != ""


def test_matrixdim_layout_matches_reference():
    # reference: CuBaseLib/cukernels.h:12-16  struct { int rows; int cols; int stride; }
    assert C.sizeof(abi.MatrixDim) == 12
    assert [f[0] for f in abi.MatrixDim._fields_] == ["rows", "cols", "stride"]
    assert C.sizeof(abi.ObjStats) == 24


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of the structs passed by pointer have the C compiler's layout of include/tnet_b200.h."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "tnet_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(TnbPeerJob), offsetof(TnbPeerJob, W), offsetof(TnbPeerJob, dW),'
                   ' offsetof(TnbPeerJob, n_frames), sizeof(TnbObjStats), sizeof(TnbGemmJob), offsetof(TnbGemmJob, A16), offsetof(TnbGemmJob, C),'
                   ' offsetof(TnbGemmJob, w_scale), offsetof(TnbGemmJob, tile_first));return 0;}\n')
    exe = str(tmp_path / "sz")
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", exe])
    got = [int(v) for v in subprocess.check_output([exe], text=True).split()]
    P = abi.PeerJob
    J = abi.GemmJob
    assert got == [C.sizeof(P), P.W.offset, P.dW.offset, P.n_frames.offset, C.sizeof(abi.ObjStats),
                   C.sizeof(J), J.A16.offset, J.C.offset, J.w_scale.offset, J.tile_first.offset]
