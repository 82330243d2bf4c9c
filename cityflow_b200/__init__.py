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


class _DeviceArray:
    """A device array of the engine exposed through ``__cuda_array_interface__`` (keeps the engine alive)."""

    def __init__(self, owner, ptr, n, typestr):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False),
                                         "version": 2, "strides": None}


def lane_observation_tensors(engine, stream=None):
    """Zero-copy per-lane observations as torch CUDA tensors.

    Returns ``(lane_ids, vehicle_count[int32], waiting_count[int32], speed_sum[float64])``: the
    arrays ``get_lane_vehicle_count`` / ``get_lane_waiting_vehicle_count`` report (engine.cpp:628-648)
    plus the per-lane sum of speeds, in ``engine.lane_ids()`` order, living on the engine's GPU.
    Nothing is copied to the host and the host does not wait: the refresh is ordered against
    ``stream`` (default: torch's current stream on the engine's device).  The tensors alias the
    engine's buffers: they are overwritten by the next call, so ``clone()`` what must survive it.
    """
    import torch
    if stream is None:
        stream = torch.cuda.current_stream(engine.device()).cuda_stream
    o = engine.observe_device(int(stream))
    n = o["n_lanes"]
    dev = torch.device("cuda", o["device"])
    mk = lambda key, ts: torch.as_tensor(_DeviceArray(engine, o[key], n, ts), device=dev)
    return (engine.lane_ids(), mk("lane_vehicle_count", "<i4"), mk("lane_waiting_count", "<i4"), mk("lane_speed_sum", "<f8"))


def set_tl_phases_tensor(engine, phases, stream=None):
    """``set_tl_phase`` for every traffic light at once from a torch CUDA tensor (no host copy).

    ``phases``: int32, contiguous, one entry per intersection in ``engine.intersection_ids()`` order
    (virtual intersections included, their entries ignored), on the engine's GPU, produced on
    ``stream`` (default: torch's current stream there).  Needs ``"rlTrafficLight": true``.
    """
    import torch
    if not (phases.is_cuda and phases.dtype == torch.int32 and phases.is_contiguous()
            and phases.device.index == engine.device() and phases.numel() == engine.num_intersections()):
        raise ValueError("phases must be a contiguous int32 CUDA tensor on the engine's device, one entry per intersection")
    if stream is None:
        stream = torch.cuda.current_stream(engine.device()).cuda_stream
    engine.set_tl_phases_device(phases.data_ptr(), int(stream))
