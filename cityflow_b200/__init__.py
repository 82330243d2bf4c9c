"""cityflow_b200 -- B200-native step engine behind the ``cityflow.Engine`` Python surface.

``Engine`` is the pybind11 class built from ``csrc/pymodule.cpp`` over the C-ABI in
``include/cityflow_b200.h`` (``libcityflow_b200.so``: host loader + sm_100a kernels).  There is no
CPU fallback: importing works anywhere (so the package can be built and inspected on a CPU box),
but constructing an ``Engine`` without a CUDA device raises ``RuntimeError``.
"""
import os as _os

_HERE = _os.path.dirname(_os.path.abspath(__file__))
LIB_PATH = _os.path.join(_HERE, "libcityflow_b200.so")

try:
    from ._cityflow_b200 import Archive, Engine, __version__, nccl_unique_id  # noqa: F401
except ImportError as _e:  # extension not built: fail loudly at use, not silently
    _IMPORT_ERROR = _e

    class Engine:  # type: ignore
        def __init__(self, *a, **k):
            raise RuntimeError(
                "cityflow_b200: the CUDA extension is not built (%s). Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` in the repo root." % (_IMPORT_ERROR,))

    __version__ = "unbuilt"

    class Archive:  # type: ignore
        def __init__(self, *a, **k):
            raise RuntimeError("cityflow_b200: the CUDA extension is not built")
