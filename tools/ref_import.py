"""Import the upstream reference (``/root/reference``) inside THIS container only.

Test/fixture-generation infrastructure: the reference is pure Python but needs ``h5py``,
``colorlog`` and ``paint`` which are absent here, so three stub modules are injected into
``sys.modules`` before ``artist`` is imported.  Nothing on
the product path, in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may import this file:
``/root/reference`` does not exist on the GPU box.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ARTIST_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "artist"))


def import_reference():
    """Return the imported reference ``artist`` package (stubs installed first)."""
    if not reference_available():
        raise ImportError(f"reference not found at {REFERENCE_ROOT}")
    if "h5py" not in sys.modules:
        h5 = types.ModuleType("h5py")
        h5.File = object
        sys.modules["h5py"] = h5
    if "colorlog" not in sys.modules:
        cl = types.ModuleType("colorlog")

        class ColoredFormatter:  # noqa: D401 - stub
            def __init__(self, *a, **k):
                pass

            def format(self, record):
                return str(record.getMessage())

        cl.ColoredFormatter = ColoredFormatter
        sys.modules["colorlog"] = cl
    if "paint" not in sys.modules:
        p = types.ModuleType("paint")
        p.__path__ = []
        pu = types.ModuleType("paint.util")
        pu.__path__ = []
        pm = types.ModuleType("paint.util.paint_mappings")
        pm.__getattr__ = lambda name: name
        sys.modules["paint"] = p
        sys.modules["paint.util"] = pu
        sys.modules["paint.util.paint_mappings"] = pm
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # importing artist.util writes ./runtime_log.txt into the CWD; keep that out of the repo
    cwd = os.getcwd()
    os.makedirs("/tmp/artist_ref_cwd", exist_ok=True)
    os.chdir("/tmp/artist_ref_cwd")
    try:
        import artist  # noqa: F401
        import artist.util  # noqa: F401
    finally:
        os.chdir(cwd)
    return sys.modules["artist"]
