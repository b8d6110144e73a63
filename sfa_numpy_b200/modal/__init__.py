"""Submodules for modal beamforming (mirror of micarray/modal/__init__.py:5-6)."""
from . import angular
from . import radial
from . import steering
