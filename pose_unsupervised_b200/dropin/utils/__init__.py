"""Overlay package: modules found here win, the rest resolve to the reference's lib/utils."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
