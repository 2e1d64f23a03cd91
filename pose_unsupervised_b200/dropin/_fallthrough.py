"""Name fall-through for the overlay modules.

An overlay module (say ``dropin/utils/transforms.py``) shadows the reference's file of the same
name, but only a few of its functions are on the lifting path.  ``reference_names`` loads the
reference's own file -- the same module name in a LATER entry of the overlay package's
``__path__`` (``pkgutil.extend_path`` put the reference's ``lib/<pkg>`` there) -- under a private
module name and returns its namespace, so that the overlay can re-export every name the
reference defines and override only the ones this repository implements.  Without the reference
on ``sys.path`` (stand-alone use), or when the reference's file cannot be imported here (for
example ``multiviews/triangulate.py`` without ``pymvg``), the result is empty and the overlay
exposes this repository's names only.
"""
import importlib.util
import os
import sys


def reference_names(package, modname, overlay_file):
    here = os.path.dirname(os.path.abspath(overlay_file))
    for d in list(getattr(package, '__path__', [])):
        if os.path.abspath(d) == here:
            continue
        cand = os.path.join(d, modname + '.py')
        if not os.path.isfile(cand):
            continue
        private = '%s._reference_%s' % (package.__name__, modname)
        mod = sys.modules.get(private)
        if mod is None:
            spec = importlib.util.spec_from_file_location(private, cand)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[private] = mod
            try:
                spec.loader.exec_module(mod)
            except ImportError:          # a third-party dependency of the reference is absent
                del sys.modules[private]
                return {}, None
        return {k: v for k, v in vars(mod).items() if not k.startswith('__')}, mod
    return {}, None
