"""Scenario file reading for the ray-tracing path: a dependency-free reader for the HDF5 subset ARTIST scenario
files use (``h5lite``), API-compatible with the ``h5py`` calls the reference's loaders make."""
from . import h5lite
from .h5lite import File

__all__ = ["File", "h5lite"]
