"""resnmtf_b200 -- B200-native ResNMTF multiplicative-update loop behind the reference's R interface
(mirrored in Python here because this image has no R).  See DESIGN.md."""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
