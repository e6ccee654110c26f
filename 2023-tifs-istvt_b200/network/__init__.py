"""Host-side mirror of the reference's `network` package for the ISTVT path (same module/attribute names,
so `state_dict` keys, `model_selection(...)` and `XceptionVidTr()` call sites keep working).

Shadow mode: with `2023-tifs-istvt_b200/` itself on `sys.path`, `import network` finds THIS directory
as a top-level package.  The modules below reach their siblings (`engine`, `ablation`, `train`, `ops`) with relative
imports that climb above `network`, so a re-rooted copy could construct a model but not run it.  Instead of re-rooting,
the top-level name is aliased to the real sub-package: the parent package is imported through importlib (repo root added
to `sys.path`) and `sys.modules['network']` plus its sub-modules are pointed at `2023-tifs-istvt_b200.network.*` — one
set of classes, whichever name they were imported under.
"""
import sys as _sys

if __name__ == "network":
    import importlib as _importlib
    import os as _os

    _pkg_dir = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    _root = _os.path.dirname(_pkg_dir)
    if _root not in _sys.path:
        _sys.path.append(_root)
    _real = _importlib.import_module(_os.path.basename(_pkg_dir) + ".network")
    for _sub in ("models", "xception", "vivit", "vivit.vivit", "vivit.module"):
        _sys.modules["network." + _sub] = _importlib.import_module(_real.__name__ + "." + _sub)
    _sys.modules["network"] = _real     # the import machinery returns sys.modules[name] after executing this file
