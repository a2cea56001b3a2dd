"""itsxpress-b200: B200-native hot path of ITSxpress behind the reference's Python API."""
from ._version import __version__  # noqa: F401
