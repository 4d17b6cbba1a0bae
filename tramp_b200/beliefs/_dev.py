"""Shared plumbing: evaluate an elementwise moment kernel on broadcast numpy /
tensor arguments and give the result back in the caller's array type."""
import numpy as np
from .. import ops


def elementwise(f, a, b, y, what):
    numpy_out = not any(ops.is_tensor(x) for x in (a, b, y) if x is not None)
    args = [np.asarray(x.cpu().numpy() if ops.is_tensor(x) else x, dtype=float)
            for x in (a, b) + ((y,) if y is not None else ())]
    args = np.broadcast_arrays(*args)
    shape = args[0].shape
    flat = [ops.padded(np.ascontiguousarray(x).reshape(1, -1)) for x in args]
    n = int(np.prod(shape)) if shape else 1
    yv = flat[2] if y is not None else None
    if what == "A":
        out = ops.factor_log_partition(f, flat[0], flat[1], yv, n, True, True)
    elif what == "p":
        out = ops.factor_posterior(f, flat[0], flat[1], yv, n, True, 2)[1]
    else:
        r, v = ops.factor_posterior(f, flat[0], flat[1], yv, n, True, True)
        out = r if what == "r" else v
    out = out[0, :n].reshape(shape)
    if numpy_out:
        out = out.cpu().numpy()
        return float(out) if not shape else out
    return out
