"""B200-native bundle-adjustment engine behind the laserCalib ``PySBA`` interface."""
from .pySBA import PySBA  # noqa: F401

__all__ = ["PySBA"]
