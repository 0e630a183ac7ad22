"""CPU tests of the boundary: the shared library loads, exports every symbol include/ols_b200.h
declares, and rejects bad arguments with the reference's messages (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

from online_lang_splatting_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "ols_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ols_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = N.lib()
    declared = _header_functions()
    assert set(declared) == set(N.EXPORTED_SYMBOLS), (declared, N.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ols_abi_version() == 2


def test_struct_layouts_match_header_sizes():
    # sizes implied by the header on LP64
    assert C.sizeof(N.RasterArgs) == 12 * 4 + 14 * 8 + 8 + 8
    assert C.sizeof(N.FwdOut) == 6 * 8
    assert C.sizeof(N.FwdInfo) == 32
    assert C.sizeof(N.BwdArgs) == 18 * 8
    assert C.sizeof(N.AEChain) == 4 + 9 * 4 + 4 + 4 + 4 + 4 + 8 * 8 * 2


def test_workspace_size_monotone_and_invalid():
    lib = N.lib()
    a = lib.ols_lang_workspace_size(1000, 15, 96, 64, 15, 10000)
    b = lib.ols_lang_workspace_size(1000, 15, 96, 64, 15, 20000)
    c = lib.ols_lang_workspace_size(2000, 15, 96, 64, 15, 10000)
    assert 0 < a < b and a < c
    assert b - a >= 12 * 10000 - 512
    assert lib.ols_lang_workspace_size(-1, 15, 96, 64, 15, 10) == 0
    assert lib.ols_lang_workspace_size(10, 15, 0, 64, 15, 10) == 0


def test_argument_validation_messages():
    lib = N.lib()
    args = N.RasterArgs(P=10, F=15, W=32, H=32, tile=15)
    out = N.FwdOut()
    rc = lib.ols_lang_forward(C.byref(args), C.byref(out), None)
    assert rc == -1
    assert b"excatly one of either SHs or precomputed colors" in lib.ols_last_error()
    args = N.RasterArgs(P=10, F=7, W=32, H=32, tile=15)
    assert lib.ols_lang_forward(C.byref(args), C.byref(out), None) == -5
    args = N.RasterArgs(P=10, F=15, W=32, H=32, tile=15, d_shs=1, M=1)
    assert lib.ols_lang_forward(C.byref(args), C.byref(out), None) == -1
    assert b"scale/rotation pair or precomputed 3D covariance" in lib.ols_last_error()


def test_python_wrapper_validation_matches_reference_messages():
    import torch
    from online_lang_splatting_b200.diff_gaussian_rasterization import (GaussianRasterizationSettings,
                                                                      LanguageGaussianRasterizer)
    rs = GaussianRasterizationSettings(32, 32, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), torch.eye(4),
                                       0, torch.zeros(3), False, False)
    assert rs.tile_size == 15 and rs.backward_mode == "compat"
    r = LanguageGaussianRasterizer(rs)
    x = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 1, 3))


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from online_lang_splatting_b200.diff_gaussian_rasterization import (GaussianRasterizationSettings,
                                                                      LanguageGaussianRasterizer)
    rs = GaussianRasterizationSettings(32, 32, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), torch.eye(4),
                                       0, torch.zeros(3), False, False)
    r = LanguageGaussianRasterizer(rs)
    x = torch.zeros(4, 3)
    with pytest.raises(RuntimeError):
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=x,
          rotations=torch.zeros(4, 4), language_precomp=torch.zeros(4, 15))


def test_header_is_plain_c_and_struct_sizes_match_ctypes(tmp_path):
    """include/ols_b200.h compiles as C99 (it is the drop-in boundary for non-C++ hosts) and every ctypes mirror in
    _native.py has exactly the size the C compiler gives the struct."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pairs = [("ols_raster_args", N.RasterArgs), ("ols_fwd_out", N.FwdOut), ("ols_fwd_info", N.FwdInfo), ("ols_bwd_args", N.BwdArgs),
             ("ols_dis_args", N.DisArgs), ("ols_dis_fwd_out", N.DisFwdOut), ("ols_dis_bwd_args", N.DisBwdArgs),
             ("ols_loss_args", N.LossArgs), ("ols_adam_group", N.AdamGroup), ("ols_ws_view", N.WsView), ("ols_host_out", N.HostOut),
             ("ols_ae_chain", N.AEChain), ("ols_hr_weights", N.HRWeights), ("ols_ssim_args", N.SsimArgs),
             ("ols_densify_params", N.DensifyParams), ("ols_pose_step", N.PoseStep)]
    src = tmp_path / "sizes.c"
    body = "\n".join('    printf("%s %%zu\\n", sizeof(%s));' % (c, c) for c, _ in pairs)
    src.write_text('#include <stdio.h>\n#include "ols_b200.h"\nint main(void) {\n%s\n    return 0;\n}\n' % body)
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, ctype in pairs:
        assert int(out[cname]) == C.sizeof(ctype), (cname, out[cname], C.sizeof(ctype))
