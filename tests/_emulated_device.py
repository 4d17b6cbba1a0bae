"""Test infrastructure: run the HOST side of tramp_b200's State Evolution on CPU
tensors, with the two C entry points it calls (`trb_se_run`, `trb_se_measure`)
emulated by the oracle evaluated with the kernels' quadrature rule.

This exists so that `-m "not gpu"` covers the Python glue around the kernels
(initialisers, damping configuration, record replay into callbacks, snapshots,
scenario / grid helpers) in the build container, which has no GPU.  It is NOT a
CPU fallback: it lives under tests/, is installed by a function-scoped fixture
and patches the loaded modules only for the duration of one test.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import se_oracle as S


def _arr(ptr, n, dtype=np.float64):
    if not ptr:
        return None
    ct = {np.float64: C.c_double, np.int32: C.c_int32}[dtype]
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,))


def _spec_of(f):
    """trb_factor -> oracle spec (inverse of tramp_b200.ops.*_factor)."""
    from scipy.special import expit
    k = f.kind
    if k == 0:
        normal_A = 0.5 * (f.p1**2 / f.p0 + np.log(2 * np.pi / f.p0))
        return dict(kind="gauss_bernoulli", rho=float(expit(normal_A - f.p2)), mean=f.p1 / f.p0,
                    var=1 / f.p0, AMIN=f.amin, AMAX=f.amax)
    if k == 1:
        return dict(kind="binary", p_pos=float(expit(2 * f.p0)), AMIN=f.amin, AMAX=f.amax)
    if k == 2:
        return dict(kind="gaussian", mean=f.p1 / f.p0, var=1 / f.p0)
    if k == 3:
        return dict(kind="gaussian", var=1 / f.p0)
    return dict(kind={4: "sgn", 5: "abs"}[k], AMIN=f.amin, AMAX=f.amax)


class EmulatedLibrary:
    """Stands in for libtramp_b200.so: everything but the SE entry points is
    forwarded to the real library (argument checks, sizes, profiling counters)."""

    def __init__(self, real, lib_module):
        self._real, self._lib = real, lib_module
        self._gl = S.Integrator("gl")
        self.calls = dict(trb_se_run=0, trb_se_measure=0)

    def __getattr__(self, name):
        return getattr(self._real, name)

    def _factors(self, ptr, n):
        size = C.sizeof(self._lib.TrbFactor)
        return [self._lib.TrbFactor.from_address(ptr + i * size) for i in range(n)]

    def trb_se_measure(self, fptr, stride, what, B, a, tau, q, out, flags, stream):
        self.calls["trb_se_measure"] += 1
        fs = self._factors(fptr, B if stride else 1)
        a_, tau_, out_, fl = _arr(a, B), _arr(tau, B), _arr(out, B), _arr(flags, B, np.int32)
        for b in range(B):
            f = fs[b * stride]
            spec = _spec_of(f)
            try:
                if f.kind <= 2:
                    fn = S.prior_forward_error if what == 0 else S.prior_free_energy
                    out_[b] = fn(spec, a_[b], self._gl)
                else:
                    fn = S.lik_backward_error if what == 0 else S.lik_free_energy
                    out_[b] = fn(spec, a_[b], tau_[b], self._gl)
            except AssertionError:
                out_[b] = np.nan
                fl[b] |= self._lib.FLAG_SE_DOMAIN
        return 0

    def trb_se_run(self, se_ref, it0, n_iter, stream):
        self.calls["trb_se_run"] += 1
        se = se_ref._obj
        G = se.G
        pri, lik = self._factors(se.prior, G), self._factors(se.lik, G)
        ea = _arr(se.edge_a, 8 * G).reshape(8, G)
        vx, vz = _arr(se.vx, G), _arr(se.vz, G)
        act, fl, ni = (_arr(p, G, np.int32) for p in (se.active, se.flags, se.n_iter))
        rvx, rvz = _arr(se.rec_vx, se.max_records * G), _arr(se.rec_vz, se.max_records * G)
        for g in range(G):
            if not act[g]:
                continue
            if se.channel == self._lib.SE_MARCHENKO_PASTUR:
                ch = dict(kind="marchenko", alpha=_arr(se.alpha, G)[g],
                          mean_spectrum=_arr(se.mean_spectrum, G)[g])
            else:
                spectrum = np.zeros(se.Nz)
                spectrum[:se.R] = _arr(se.s2, se.R)
                ch = dict(kind="spectrum", spectrum=spectrum, Nx=se.Nx, rank=se.rank)
            ch.update(AMIN=se.lin_amin, AMAX=se.lin_amax)
            early = None
            if se.es_tol >= 0:
                ids = tuple(k for k, bit in (("x", 1), ("z", 2)) if se.es_vars & bit)
                early = dict(tol=se.es_tol, min_variance=se.es_min_variance,
                             wait_increase=se.es_wait_increase, max_increase=se.es_max_increase, ids=ids)
            try:
                r = S.se_glm(_spec_of(pri[g]), ch, _spec_of(lik[g]), n_iter,
                             dict(e1=se.damp1, e3=se.damp3, e5=se.damp5, e7=se.damp7),
                             {f"e{k + 1}": ea[k, g] for k in range(8)}, early, self._gl)
            except AssertionError:
                fl[g] |= self._lib.FLAG_SE_DOMAIN
                act[g] = 0
                continue
            ea[:, g] = r["a"]
            vx[g], vz[g] = r["v"]
            ni[g] += r["n_iter"]
            if rvx is not None:
                for i in range(r["n_iter"]):
                    rvx[(it0 + i) * G + g] = r["vx"][i]
                    rvz[(it0 + i) * G + g] = r["vz"][i]
            if r["n_iter"] < n_iter:
                act[g] = 0
        return 0


def install(setattr_):
    """Patch the loaded tramp_b200 modules through `setattr_(obj, name, value)`
    (pytest's monkeypatch.setattr in the fixture, plain setattr in a spawned
    worker process of a multi-rank test)."""
    from tramp_b200 import _lib, ops
    fake = EmulatedLibrary(_lib.load(), _lib)
    setattr_(ops, "device", lambda: torch.device("cpu"))
    setattr_(ops, "_quad_cache", {})
    setattr_(_lib, "require_cuda", lambda: None)
    setattr_(_lib, "current_stream", lambda: None)
    setattr_(ops, "current_stream", lambda: None)
    setattr_(_lib, "load", lambda: fake)
    return fake


@pytest.fixture
def emulated_device(monkeypatch):
    """tramp_b200 with CPU tensors and the SE kernels emulated by the oracle."""
    return install(monkeypatch.setattr)
