"""Truncated-normal mean / variance / log-partition / probability
(reference tramp/utils/truncated_normal.py:234-298) on the GPU: the erfcx
half-infinite fast path and the five-branch finite-interval F0/F1/F2 live in
tramp_b200/csrc/trb_moments.cuh (`truncated_normal`)."""
import numpy as np
from .. import ops


def _run(r0, v0, zmin, zmax, which):
    assert zmin < zmax
    numpy_out = not (ops.is_tensor(r0) or ops.is_tensor(v0))
    if numpy_out:
        r0_, v0_ = np.broadcast_arrays(np.asarray(r0, float), np.asarray(v0, float))
        shape = r0_.shape
        r0_d, v0_d = ops.to_dev(np.ascontiguousarray(r0_).reshape(-1)), \
            ops.to_dev(np.ascontiguousarray(v0_).reshape(-1))
    else:
        t = ops.torch()
        r0_d, v0_d = t.broadcast_tensors(ops.to_dev(r0), ops.to_dev(v0))
        shape = tuple(r0_d.shape)
        r0_d, v0_d = r0_d.reshape(-1).contiguous(), v0_d.reshape(-1).contiguous()
    outs = ops.truncated_normal(r0_d, v0_d, zmin, zmax)

    def give(out):
        out = out.reshape(shape)
        if numpy_out:
            out = out.cpu().numpy()
            return float(out) if not shape else out
        return out
    if which is None:
        return tuple(give(o) for o in outs)
    return give(outs[which])


def truncated_normal_moments(r0, v0, zmin, zmax, only=None):
    """(mean, var, logZ, proba) from ONE launch; `only` = 0..3 returns that one."""
    return _run(r0, v0, zmin, zmax, only)


def truncated_normal_mean(r0, v0, zmin, zmax):
    "Mean of N(z | r0 v0) restricted to [zmin, zmax]"
    return _run(r0, v0, zmin, zmax, 0)


def truncated_normal_var(r0, v0, zmin, zmax):
    "Variance of N(z | r0 v0) restricted to [zmin, zmax]"
    return _run(r0, v0, zmin, zmax, 1)


def truncated_normal_logZ(r0, v0, zmin, zmax):
    "Log partition of N(z | r0 v0) restricted to [zmin, zmax]"
    return _run(r0, v0, zmin, zmax, 2)


def truncated_normal_proba(r0, v0, zmin, zmax):
    "Probability that z ~ N(r0, v0) falls in [zmin, zmax]"
    return _run(r0, v0, zmin, zmax, 3)
