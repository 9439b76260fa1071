"""Importable alias of the product package, whose directory is named `monkey-pose_b200/` (a hyphen
cannot appear in a Python module name).  All code lives there; this file only points at it."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..",
                                 "monkey-pose_b200"))

from ._api import *  # noqa: F401,F403,E402
from ._api import __all__  # noqa: F401,E402
