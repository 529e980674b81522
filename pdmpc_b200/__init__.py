"""Importable alias of the ``p-dmpc_b200/`` package directory.

The product lives in ``p-dmpc_b200/`` (the name the build contract asks for);
a hyphen is not a valid Python identifier, so this stub package extends its
``__path__`` to that directory.  ``import pdmpc_b200.capi`` therefore loads
``p-dmpc_b200/capi.py``.
"""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
__path__.append(_os.path.normpath(_os.path.join(_here, "..", "p-dmpc_b200")))
