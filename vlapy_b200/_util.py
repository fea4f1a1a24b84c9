"""Host <-> device plumbing shared by the operator mirrors (torch is memory + streams only)."""
import numpy as np
import torch


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("vlapy_b200: the b200 backend needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def to_dev(x):
    """numpy / scalar / tensor -> (contiguous CUDA fp64 tensor, came_from_host)."""
    if isinstance(x, torch.Tensor):
        if x.is_cuda and x.dtype == torch.float64:
            return x, False
        return x.to(device=device(), dtype=torch.float64), not x.is_cuda
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    return torch.from_numpy(a).to(device()), True


def const(x):
    """upload a static host array once (wavenumbers, grids)"""
    return to_dev(x)[0].contiguous()


def back(t, to_host):
    """return the same kind the caller passed in"""
    return t.cpu().numpy() if to_host else t


def attach_density(f, n):
    """remember the charge density n = trapz_v f that the producing kernel already reduced, together with
    the tensor version it belongs to: an in-place change of f afterwards invalidates it"""
    f._vpfp_density = (n, f._version)
    return f


def cached_density(f):
    """the density attached by ``attach_density`` if f has not been modified in place since, else None
    (the caller then integrates f itself, as vlapy/core/field.py:27-36 always does)"""
    c = getattr(f, "_vpfp_density", None)
    if c is None or c[1] != f._version:
        return None
    return c[0]
