"""cffi (ABI mode) binding of libasp_b200.so -- the thin C-ABI layer, in the style of the
reference's annealing_sign_problem/build_extension.py.  No CPU fallback: a missing library
or a missing CUDA device raises."""
import os

import torch

from .build_extension import LIBRARY, ffibuilder

ffi = ffibuilder
_lib = None


class AspError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIBRARY):
            raise AspError(
                "libasp_b200.so is not built (run `python annealing-sign-problem_b200/build_extension.py`); "
                "this package has no CPU fallback")
        # ASP_B200_LIBRARY: development override (kernel variants built side by side)
        _lib = ffi.dlopen(os.environ.get("ASP_B200_LIBRARY", LIBRARY))
    return _lib


def require_cuda() -> torch.device:
    if not torch.cuda.is_available() or lib().asp_device_count() < 1:
        raise AspError("no CUDA device: the B200 hot path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def check(rc: int) -> None:
    if rc != 0:
        raise AspError("asp_b200 error %d: %s" % (rc, ffi.string(lib().asp_last_error()).decode()))


def ptr(tensor, ctype: str):
    """Device (or host) pointer of a contiguous torch tensor as a cffi pointer."""
    if tensor is None:
        return ffi.NULL
    assert tensor.is_contiguous()
    return ffi.cast(ctype, tensor.data_ptr())


def stream():
    return ffi.cast("void *", torch.cuda.current_stream().cuda_stream)


def as_u64(t: torch.Tensor) -> torch.Tensor:
    """torch has no general uint64 arithmetic: basis words travel as int64 bit patterns."""
    return t.view(torch.int64) if t.dtype == torch.uint64 else t
