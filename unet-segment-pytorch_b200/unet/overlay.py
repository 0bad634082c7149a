"""Fall-through from this package to the reference's ``unet`` package.

This package owns only the hot path (``unet.models``, ``unet.utils.loss``, ``unet.utils.metrics``,
``ModelEMA``, the device-side ``unet.data`` pipeline).  The reference's scripts also import
orchestration modules that are out of scope to rebuild (scripts/train.py:28-35,
scripts/predict.py:29-30): ``unet.data.dataset``, ``unet.data.augmentations``,
``unet.utils.callbacks``, ``unet.utils.plots`` and the helpers of ``unet.utils.general``
(``set_seed``, ``get_device``, ``load_config``, ``increment_path``).  With a reference checkout
attached, those resolve to the reference's own files, unchanged:

* sub-MODULES this package does not have are found through ``__path__``: the reference's
  ``unet/``, ``unet/utils/``, ``unet/data/`` directories are appended to the search path of the
  packages of the same name here (this package's files always win);
* NAMES missing from a module that exists on both sides (``unet.utils.general``, the package
  ``__init__`` files) are looked up, on first use, in the reference's file of the same name, loaded
  under a private module name (``unet.utils._reference_general`` …).

The reference root is ``attach(root)``'s argument, else ``$UNET_REFERENCE_ROOT``, else the first
``sys.path`` entry that holds another ``unet/__init__.py``.  ``python -m unet.dropin
<reference>/scripts/train.py …`` attaches the script's checkout and runs it (see ``dropin.py``).
"""
from __future__ import annotations

import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = None          # attached reference checkout (directory that holds unet/)
_SHADOW = {}          # our module name -> loaded reference twin

# packages of this distribution whose search path is extended: package name -> sub-directory of unet/
_PACKAGES = {"unet": "", "unet.utils": "utils", "unet.data": "data", "unet.models": "models"}


def _is_reference_root(path: str) -> bool:
    init = os.path.join(path, "unet", "__init__.py")
    return os.path.isfile(init) and os.path.realpath(os.path.join(path, "unet")) != os.path.realpath(_HERE)


def find_root():
    """The reference checkout to fall through to, or None."""
    env = os.environ.get("UNET_REFERENCE_ROOT")
    if env:
        if not _is_reference_root(env):
            raise RuntimeError(f"UNET_REFERENCE_ROOT={env!r} does not hold the reference's unet/ package")
        return os.path.abspath(env)
    for p in sys.path:
        if p and os.path.isdir(p) and _is_reference_root(p):
            return os.path.abspath(p)
    return None


def root():
    """Attached reference root (attaching lazily from the environment / sys.path), or None."""
    if _ROOT is None:
        r = find_root()
        if r is not None:
            attach(r)
    return _ROOT


def attach(reference_root: str) -> str:
    """Make the reference checkout at ``reference_root`` the fall-through target."""
    global _ROOT
    reference_root = os.path.abspath(reference_root)
    if not _is_reference_root(reference_root):
        raise RuntimeError(f"{reference_root!r} does not hold the reference's unet/ package")
    if _ROOT is not None and _ROOT != reference_root:
        raise RuntimeError(f"a different reference checkout is already attached: {_ROOT}")
    _ROOT = reference_root
    for name in _PACKAGES:
        mod = sys.modules.get(name)
        if mod is not None:
            extend_path(mod)
    return _ROOT


def extend_path(package) -> None:
    """Append the reference's directory of the same package to ``package.__path__`` (ours first)."""
    if _ROOT is None:
        return
    sub = _PACKAGES.get(package.__name__)
    if sub is None:
        return
    d = os.path.join(_ROOT, "unet", sub) if sub else os.path.join(_ROOT, "unet")
    if os.path.isdir(d) and d not in package.__path__:
        package.__path__.append(d)


def _reference_twin(module_name: str, file_rel: str):
    """The reference's file ``unet/<file_rel>`` loaded as a private sibling module (so that its
    relative imports keep working) — or None without an attached checkout / without that file."""
    if module_name in _SHADOW:
        return _SHADOW[module_name]
    if root() is None:
        return None
    path = os.path.join(_ROOT, "unet", file_rel)
    if not os.path.isfile(path):
        return None
    pkg, _, leaf = module_name.rpartition(".")
    is_pkg = os.path.basename(path) == "__init__.py"
    if is_pkg:
        return None   # package __init__ twins are not executed: their names come from submodules
    private = f"{pkg}._reference_{leaf}"
    spec = importlib.util.spec_from_file_location(private, path)
    twin = importlib.util.module_from_spec(spec)
    sys.modules[private] = twin
    try:
        spec.loader.exec_module(twin)
    except BaseException:
        sys.modules.pop(private, None)
        raise
    _SHADOW[module_name] = twin
    return twin


def module_getattr(module_name: str, file_rel: str):
    """``__getattr__`` for a module that exists on both sides: names this package does not define
    come from the reference's file of the same name."""

    def __getattr__(name: str):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        twin = _reference_twin(module_name, file_rel)
        if twin is not None and hasattr(twin, name):
            return getattr(twin, name)
        hint = "" if _ROOT is not None else (" (no reference checkout attached: set UNET_REFERENCE_ROOT or call "
                                              "unet.overlay.attach(root) to fall through to the reference's module)")
        raise AttributeError(f"module {module_name!r} has no attribute {name!r}{hint}")

    return __getattr__


def package_getattr(package_name: str, submodules):
    """``__getattr__`` for a package ``__init__``: a missing name is looked up in the listed
    submodules (ours or, through the extended ``__path__``, the reference's) — this is how
    ``from unet.utils import set_seed, EarlyStopping`` (the reference's ``unet/utils/__init__.py``
    re-exports) keeps working."""

    def __getattr__(name: str):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        import importlib

        pkg = sys.modules[package_name]
        root()
        extend_path(pkg)
        try:   # a submodule of that name (unet.utils.callbacks, unet.data.dataset, …)
            return importlib.import_module(f"{package_name}.{name}")
        except ModuleNotFoundError as e:
            if e.name != f"{package_name}.{name}":
                raise
        for sub in submodules:
            try:
                mod = importlib.import_module(f"{package_name}.{sub}")
            except ModuleNotFoundError as e:
                if e.name != f"{package_name}.{sub}":
                    raise
                continue
            try:
                return getattr(mod, name)
            except AttributeError:
                continue
        raise AttributeError(f"module {package_name!r} has no attribute {name!r}")

    return __getattr__


class _LateFinder:
    """Last entry of ``sys.meta_path``: a ``unet.*`` submodule nobody found yet.  The reference
    checkout may have been put on ``sys.path`` after this package was imported (the reference's
    scripts insert their own root at run time): look again, extend the search paths, retry."""

    @staticmethod
    def find_spec(fullname, path=None, target=None):
        parent, _, _ = fullname.rpartition(".")
        pkg = sys.modules.get(parent)
        if parent not in _PACKAGES or pkg is None or root() is None:
            return None
        extend_path(pkg)
        from importlib.machinery import PathFinder

        # only the reference's directory: this package's own directories were searched already
        return PathFinder.find_spec(fullname, [d for d in pkg.__path__ if d.startswith(_ROOT + os.sep)])


def install(package) -> None:
    """Called by each package ``__init__`` of this distribution."""
    if not any(isinstance(f, type) and f.__name__ == "_LateFinder" for f in sys.meta_path):
        sys.meta_path.append(_LateFinder)
    if _ROOT is not None or os.environ.get("UNET_REFERENCE_ROOT"):
        root()
    extend_path(package)
