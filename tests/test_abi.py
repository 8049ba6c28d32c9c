"""CPU: the drop-in boundary.  The C-ABI library loads and exports every symbol include/gof_rasterizer.h
declares; the Python package has the reference's surface (names, field order, argument validation); the
product refuses to run without CUDA instead of falling back."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gof_rasterizer.h")).read()
    return sorted(set(re.findall(r"GOF_API\s+[\w\s\*]+?\b(gof_\w+)\s*\(", text)))


def test_header_declares_the_reference_entry_points():
    syms = _declared_symbols()
    for s in ("gof_rasterize_forward", "gof_rasterize_backward", "gof_integrate", "gof_mark_visible",
              "gof_export_state", "gof_marching_tets_count", "gof_marching_tets_emit", "gof_last_error", "gof_version"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    lib_path = os.path.join(ROOT, "gaussian-opacity-fields_b200", "diff_gaussian_rasterization", "libgof_b200.so")
    assert os.path.exists(lib_path), "build the library first: python gaussian-opacity-fields_b200/build.py"
    lib = ctypes.CDLL(lib_path)
    for s in _declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/gof_rasterizer.h but not exported"
    lib.gof_version.restype = ctypes.c_int
    assert lib.gof_version() >= 100


def test_library_has_no_torch_or_python_dependency():
    import subprocess
    lib_path = os.path.join(ROOT, "gaussian-opacity-fields_b200", "diff_gaussian_rasterization", "libgof_b200.so")
    out = subprocess.run(["ldd", lib_path], capture_output=True, text=True).stdout
    names = [line.split()[0] for line in out.splitlines() if line.strip()]
    assert not any(("torch" in n or "python" in n or "c10" in n) for n in names), names


def test_python_surface_matches_reference():
    import diff_gaussian_rasterization as dgr
    assert dgr.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "kernel_size", "subpixel_offset", "bg", "scale_modifier",
        "viewmatrix", "projmatrix", "sh_degree", "campos", "prefiltered", "debug")
    fwd = inspect.signature(dgr.GaussianRasterizer.forward)
    assert list(fwd.parameters)[1:] == ["means3D", "means2D", "opacities", "shs", "colors_precomp", "scales", "rotations",
                                        "cov3D_precomp", "view2gaussian_precomp"]
    integ = inspect.signature(dgr.GaussianRasterizer.integrate)
    assert list(integ.parameters)[1:] == ["points3D", "means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                          "rotations", "cov3D_precomp", "view2gaussian_precomp"]
    rg = inspect.signature(dgr.rasterize_gaussians)
    assert list(rg.parameters) == ["means3D", "means2D", "sh", "colors_precomp", "opacities", "scales", "rotations",
                                   "cov3Ds_precomp", "view2gaussian_precomp", "raster_settings"]
    for name in ("rasterize_gaussians", "rasterize_gaussians_backward", "integrate_gaussians_to_points", "mark_visible"):
        assert callable(getattr(dgr._C, name))
    assert hasattr(dgr.GaussianRasterizer, "markVisible")


def _settings(dgr, H=32, W=32):
    z = torch.zeros
    return dgr.GaussianRasterizationSettings(H, W, 0.5, 0.5, 0.0, z(H, W, 2), z(3), 1.0, torch.eye(4), torch.eye(4), 3, z(3),
                                             False, False)


def test_argument_validation_mirrors_reference():
    import diff_gaussian_rasterization as dgr
    r = dgr.GaussianRasterizer(_settings(dgr))
    P = 4
    m, o = torch.zeros(P, 3), torch.zeros(P, 1)
    sh, col, sc, rot, cov = torch.zeros(P, 16, 3), torch.zeros(P, 3), torch.ones(P, 3), torch.zeros(P, 4), torch.zeros(P, 6)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(m, m, o, shs=None, colors_precomp=None, scales=sc, rotations=rot)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(m, m, o, shs=sh, colors_precomp=col, scales=sc, rotations=rot)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(m, m, o, shs=sh, scales=sc, rotations=None)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(m, m, o, shs=sh, scales=sc, rotations=rot, cov3D_precomp=cov)


def test_no_cpu_fallback():
    """CPU tensors must be rejected loudly -- there is no CPU or PyTorch path in the product."""
    import diff_gaussian_rasterization as dgr
    r = dgr.GaussianRasterizer(_settings(dgr))
    P = 4
    m, o = torch.zeros(P, 3), torch.zeros(P, 1)
    with pytest.raises(Exception):
        r(m, m, o, shs=torch.zeros(P, 16, 3), scales=torch.ones(P, 3), rotations=torch.zeros(P, 4))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "gaussian-opacity-fields_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "gof_oracle" not in text and "oracle/" not in text.replace("the oracle", ""), f"{f} references oracle/"


def test_scratch_buffers_die_by_refcount_not_by_gc():
    """The allocator thunk must not form a reference cycle with the buffer it hands out: the scratch tensor has to
    be released (to the pool / torch's caching allocator) the moment the caller drops it."""
    import gc
    import weakref
    from diff_gaussian_rasterization import _C
    gc.disable()
    try:
        sc = _C._Scratch(torch.device("cpu"))
        assert sc.cb(None, 4096) != 0
        w = weakref.ref(sc.tensor)
        assert w() is not None and w().numel() == 4096
        del sc
        assert w() is None, "scratch tensor kept alive by a reference cycle"
    finally:
        gc.enable()


def test_header_is_plain_c(tmp_path):
    """include/gof_rasterizer.h is the C-ABI contract: it must compile as C99 (no torch, no C++) and as C++."""
    import subprocess
    inc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include")
    c = tmp_path / "h.c"
    c.write_text('#include "gof_rasterizer.h"\nint main(void) { return gof_version() == 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, "-c", str(c), "-o", str(tmp_path / "h.o")])
    cpp = tmp_path / "h.cpp"
    cpp.write_text('#include "gof_rasterizer.h"\nint main() { return gof_version() == 0; }\n')
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Werror", "-I", inc, "-c", str(cpp), "-o", str(tmp_path / "h2.o")])
