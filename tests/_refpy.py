"""TEST INFRASTRUCTURE: loads the reference's UNMODIFIED Python (staged by baseline/stage_ref.sh into the git-ignored
baseline/_ref/gof_ref_py/) under alias module names, so that it can run side by side with this repo's package of the same
name.  Nothing here is product code; every loader returns None when the staged files (or oracle/_ref) are absent."""
import ast
import importlib.util
import os
import sys
import types

import _util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = os.path.join(ROOT, "baseline", "_ref", "gof_ref_py")


def staged(*parts):
    p = os.path.join(STAGE, *parts)
    return p if os.path.exists(p) else None


def _load(alias, path, package=False):
    spec = importlib.util.spec_from_file_location(alias, path, submodule_search_locations=[os.path.dirname(path)] if package else None)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


_CACHE = {}


def ref_rasterizer_package():
    """The reference's `diff_gaussian_rasterization` Python package (its autograd Function and GaussianRasterizer module,
    DGR/diff_gaussian_rasterization/__init__.py) on top of the reference's compiled `_C` (oracle/_ref), as module `gof_ref_dgr`."""
    if "dgr" in _CACHE:
        return _CACHE["dgr"]
    path, refc = staged("diff_gaussian_rasterization", "__init__.py"), _util.load_ref()
    if path is None or refc is None:
        return None
    sys.modules["gof_ref_dgr._C"] = refc              # `from . import _C` inside the package resolves to the compiled reference
    _CACHE["dgr"] = _load("gof_ref_dgr", path, package=True)
    return _CACHE["dgr"]


def ref_utils(name):
    """utils/<name>.py of the reference (sh_utils, loss_utils, depth_utils, general_utils, graphics_utils, tetmesh)."""
    key = "utils." + name
    if key not in _CACHE:
        path = staged("utils", name + ".py")
        if path is None:
            return None
        if "utils" not in sys.modules or not hasattr(sys.modules["utils"], "__gof_ref__"):
            pkg = types.ModuleType("utils")
            pkg.__path__ = [os.path.join(STAGE, "utils")]
            pkg.__gof_ref__ = True
            sys.modules["utils"] = pkg
        _CACHE[key] = _load("utils." + name, path)
        setattr(sys.modules["utils"], name, _CACHE[key])
    return _CACHE[key]


def ref_gaussian_renderer(rasterizer_module, alias):
    """The reference's gaussian_renderer (render / integrate, gaussian_renderer/__init__.py:18-218) bound to
    `rasterizer_module` as its `diff_gaussian_rasterization` -- ours or the reference's.  `scene.gaussian_model.GaussianModel`
    is only a type annotation there; a stub class satisfies the import."""
    path = staged("gaussian_renderer", "__init__.py")
    if path is None or ref_utils("sh_utils") is None:
        return None
    saved = {k: sys.modules.get(k) for k in ("diff_gaussian_rasterization", "scene", "scene.gaussian_model")}
    scene = types.ModuleType("scene")
    scene.__path__ = []
    gm = types.ModuleType("scene.gaussian_model")
    gm.GaussianModel = type("GaussianModel", (), {})
    scene.gaussian_model = gm
    sys.modules.update({"diff_gaussian_rasterization": rasterizer_module, "scene": scene, "scene.gaussian_model": gm})
    try:
        return _load(alias, path)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def ref_function(text_file, name, glb):
    """One top-level function (or class) of a reference script that cannot be imported here (train.py, extract_mesh.py,
    scene/gaussian_model.py import packages this image lacks), compiled from its unmodified source text into `glb`."""
    path = staged("text", text_file)
    if path is None:
        return None
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name == name:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, glb)
            return glb[name]
    raise KeyError(f"{name} not found in {text_file}")


def ref_method_source(text_file, cls, name):
    """Source text of method `cls.name` (dedented) from a staged reference script."""
    import textwrap
    path = staged("text", text_file)
    if path is None:
        return None
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == name:
                    return textwrap.dedent(ast.get_source_segment(src, sub))
    raise KeyError(f"{cls}.{name} not found in {text_file}")
