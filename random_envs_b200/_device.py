"""Device plumbing shared by the env classes: torch owns memory and streams, nothing else."""
import ctypes

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda(device=None):
    """Return a torch.device for the GPU or raise: the package has no CPU code path."""
    t = torch()
    if not t.cuda.is_available():
        raise RuntimeError("random_envs_b200 needs a CUDA device (NVIDIA B200, sm_100a): "
                           "there is no CPU fallback for step/reset/sample")
    dev = t.device(device) if device is not None else t.device("cuda", t.cuda.current_device())
    if dev.type != "cuda":
        raise RuntimeError("random_envs_b200 buffers must live on a CUDA device, got %s" % dev)
    if dev.index is None:
        dev = t.device("cuda", t.cuda.current_device())
    return dev


def stream_ptr(device):
    return ctypes.c_void_p(raw_stream(device.index))


def raw_stream(device_index):
    """cudaStream_t of torch's current stream on that device, as an int.  torch._C._cuda_getCurrentRawStream skips
    the Stream object (0.3 us instead of 1.5 us per call -- it matters for 10 us kernels); falls back if it is gone."""
    t = torch()
    fast = getattr(t._C, "_cuda_getCurrentRawStream", None)
    if fast is not None:
        return fast(device_index)
    return t.cuda.current_stream(device_index).cuda_stream


def ptr(tensor):
    return ctypes.c_void_p(tensor.data_ptr()) if tensor is not None else None
