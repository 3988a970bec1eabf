"""ctypes loader for libvpf_b200.so -- the only door from Python to the kernels.

There is no fallback of any kind: if the library is missing, or an entry
returns an error, a RuntimeError is raised.  Entry signatures are declared in
include/vpf.h.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libvpf_b200.so")
_lib = None


class VpfError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise VpfError(
                f"{SO_PATH} not found: build it with `python -m vipformer_b200.build` "
                "(vipformer_b200 has no CPU / PyTorch fallback)")
        _lib = ctypes.CDLL(SO_PATH)
        _lib.vpf_last_error_string.restype = ctypes.c_char_p
        _lib.vpf_launch_count.restype = ctypes.c_int64
        _lib.vpf_abi_version.restype = ctypes.c_int
    return _lib


def call(name, *args):
    """Call an `int vpf_*(...)` entry; raise VpfError on a non-zero return."""
    fn = getattr(lib(), name)
    rc = fn(*args)
    if rc != 0:
        raise VpfError(f"{name} failed ({rc}): {lib().vpf_last_error_string().decode()}")


def size_query(name, *args):
    fn = getattr(lib(), name)
    fn.restype = ctypes.c_size_t
    return fn(*args)


def ptr(t):
    """Device (or host) pointer of a torch tensor / None -> c_void_p."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    """Raw handle of torch's current CUDA stream on the current device.  Called once per kernel launch, so it uses the
    two C entry points underneath torch.cuda.current_stream() (7 us -> well under 1 us per call)."""
    import torch

    try:
        return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
    except AttributeError:      # private names moved: the public (slower) route
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(lib().vpf_launch_count())


def launch_count_reset():
    lib().vpf_launch_count_reset()


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise VpfError("vipformer_b200 kernels need CUDA tensors (there is no CPU path); got a "
                           f"{t.device} tensor")
