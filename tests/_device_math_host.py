"""Test infrastructure: the device routines of tramp_b200/csrc/trb_moments.cuh
(prior / likelihood moments, log-partitions, truncated normal) compiled as HOST
functions, so that the formulas the kernels evaluate can be exercised in the build
container, which has no GPU.

How: the text of the header, minus its include of the device helpers, is compiled by
nvcc with `__device__` defined away; nvcc's host math library supplies erfcx.  Host and
device special functions differ by an ulp or two.  Used by
tests/test_device_math_on_host.py (the formulas against the golden vectors) and by
tests/_emulated_device.py (the factor primitives of the emulated library).  Nothing in
the package can load this library.
"""
import ctypes as C
import functools
import os
import re
import shutil
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "tramp_b200", "csrc")

WRAPPERS = r"""
extern "C" {
void hm_truncated_normal(long n, const double* r0, const double* v0, double zmin, double zmax,
                         double* mean, double* var, double* logZ, double* proba) {
  for (long i = 0; i < n; ++i) {
    const trb::TruncMoments t = trb::truncated_normal(r0[i], v0[i], zmin, zmax);
    mean[i] = t.mean; var[i] = t.var; logZ[i] = t.logZ; proba[i] = t.proba;
  }
}
// a_stride / y_stride: 1 = one value per element, 0 = a[0] for all (y == NULL: y = 0)
void hm_factor(const trb_factor* f, long n, const double* a, long a_stride, const double* b,
               const double* y, double* r, double* v, double* logZ) {
  for (long i = 0; i < n; ++i) {
    const double yi = y ? y[i] : 0.0, ai = a[i * a_stride];
    const trb::RV o = trb::factor_moments(*f, ai, b[i], yi);
    r[i] = o.r; v[i] = o.v;
    logZ[i] = trb::factor_log_partition(*f, ai, b[i], yi);
  }
}
void hm_sparse_weight(const trb_factor* f, long n, const double* a, long a_stride, const double* b, double* p) {
  for (long i = 0; i < n; ++i) p[i] = trb::sparse_weight(*f, a[i * a_stride], b[i]);
}
int hm_is_constant_message(int kind) { return trb::factor_is_constant_message(kind) ? 1 : 0; }
}
"""


def nvcc_path():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return nvcc if os.path.exists(nvcc) else None


@functools.lru_cache(maxsize=1)
def load():
    """Build (once per process, in a temporary directory) and load the library;
    None if nvcc is not available."""
    nvcc = nvcc_path()
    if nvcc is None:
        return None
    header = open(os.path.join(CSRC, "trb_moments.cuh")).read()
    assert '#include "trb_common.cuh"' in header
    body = header.replace("#pragma once", "").replace('#include "trb_common.cuh"', "")
    two_pi = re.search(r"constexpr double kTwoPi = [^;]+;", open(os.path.join(CSRC, "trb_common.cuh")).read())
    assert two_pi, "kTwoPi moved out of trb_common.cuh"
    src = "\n".join([
        "#include <cuda_runtime.h>", "#include <math.h>", '#include "tramp_b200.h"',
        "#undef __device__", "#define __device__", "#undef __forceinline__", "#define __forceinline__ inline",
        "namespace trb { " + two_pi.group(0) + " }", body, WRAPPERS])
    d = tempfile.mkdtemp(prefix="trb_host_math_")
    cu, so = os.path.join(d, "moments_host.cu"), os.path.join(d, "libmoments_host.so")
    with open(cu, "w") as fh:
        fh.write(src)
    subprocess.run([nvcc, "-O2", "-shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
                    "-I", os.path.join(ROOT, "include"), "-o", so, cu], check=True, capture_output=True)
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    lib.hm_truncated_normal.argtypes = [C.c_long, dp, dp, C.c_double, C.c_double, dp, dp, dp, dp]
    lib.hm_factor.argtypes = [C.c_void_p, C.c_long, dp, C.c_long, dp, dp, dp, dp, dp]
    lib.hm_sparse_weight.argtypes = [C.c_void_p, C.c_long, dp, C.c_long, dp, dp]
    return lib


def _p(x):
    return None if x is None else x.ctypes.data_as(C.POINTER(C.c_double))


def factor_elementwise(f, a, b, y=None):
    """(r, v, logZ) elementwise for the TrbFactor f; a scalar or array like b."""
    import numpy as np
    b = np.ascontiguousarray(b, dtype=float)
    a = np.ascontiguousarray(a, dtype=float)
    y = None if y is None else np.ascontiguousarray(y, dtype=float)
    a_stride = 0 if a.size == 1 and b.size != 1 else 1
    assert a_stride == 0 or a.shape == b.shape
    r, v, A = (np.empty_like(b) for _ in range(3))
    load().hm_factor(C.addressof(f), b.size, _p(a.reshape(-1)), a_stride, _p(b), _p(y), _p(r), _p(v), _p(A))
    return r, v, A


def sparse_weight(f, a, b):
    """beliefs/sparse.py `p` elementwise through the device routine."""
    import numpy as np
    b = np.ascontiguousarray(b, dtype=float)
    a = np.ascontiguousarray(a, dtype=float)
    a_stride = 0 if a.size == 1 and b.size != 1 else 1
    out = np.empty_like(b)
    load().hm_sparse_weight(C.addressof(f), b.size, _p(a.reshape(-1)), a_stride, _p(b), _p(out))
    return out
