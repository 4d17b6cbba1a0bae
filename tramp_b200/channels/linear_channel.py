"""LinearChannel x = W z with the Gaussian posterior diagonalised by a THIN SVD.

reference: channels/linear/linear_channel.py.  The reference keeps the full
U (Nx x Nx), V (Nz x Nz) and a dense rectangular S and applies nine dense
GEMVs per EP iteration.  Here W = U_R diag(s) V_R^T with R = min(Nx, Nz); the
operators live on the GPU as rows of singular vectors (`Vt[B, R, ldn]`,
`Ut[B, R, ldm]`) and every mean is  project -> spectrum rescale -> expand
(tramp_b200/csrc/trb_linear.cu).  With bz fixed inside an iteration this
streams each operator once per use.
"""
import logging
import numpy as np

from .base_channel import Channel
from ..base import _Arg
from .. import ops

logger = logging.getLogger(__name__)


def _ceil_to(n, m):
    return -(-int(n) // m) * m


# what the last hand-written factorisation did (route, sweeps, convergence history): read by the
# benchmark to report the FP64 rate of the set-up
LAST_SETUP_STATS = {}

# sweeps of the 32 x 32 Jacobi that diagonalises a block pair's Gram matrix: the outer iteration
# needs the same number of sweeps with 1 as with a fully converged inner solve (measured)
JACOBI_INNER_SWEEPS = 1


def jacobi_orthogonalise_rows(A, tol=1e-13, max_sweeps=40, max_inner=None):
    """Rows of A[b] <- Q_b^T A[b] (Q_b orthogonal) until the rows of every instance are mutually
    orthogonal: one-sided block Jacobi, `trb_jacobi_sweep` (tramp_b200/csrc/trb_setup.cu).

    A [B, n_rows, ld] with n_rows % 32 == 0, ld % 64 == 0, zero padded; rotated in place.
    A sweep reports the largest cosine between two rows it met BEFORE rotating them; the
    iteration stops when that is below `tol` (or has stalled at the rounding floor).  One
    read-back per sweep.  Returns the list of the sweeps' measures."""
    t = ops.torch()
    B, n_rows, ld = A.shape
    work = ops.jacobi_workspace(B, n_rows, ld, A.device)
    max_inner = max_inner or JACOBI_INNER_SWEEPS
    history = []
    for sweep in range(max_sweeps):
        off = float(ops.jacobi_sweep(A, work, skip_tol=0.05 * tol, max_inner=max_inner).max().item())
        history.append(off)
        if off != off:
            raise FloatingPointError("block Jacobi: NaN in the matrix being factorised")
        if off < tol or (len(history) > 1 and off < 1e-10 and off > 0.5 * history[-2]):
            break
    else:
        logger.warning(f"block Jacobi: largest cosine {history[-1]:.1e} after {max_sweeps} sweeps")
    return history


def _gram_rows(Wb, out):
    """out[:M, :M] = Wb Wb^T for Wb [M, N] (DMMA GEMM of trb_gemm.cu)."""
    M, N = Wb.shape
    if out.shape[-1] == M and out.is_contiguous():
        ops.lin_project_gemm(Wb, M, N, Wb, M, out=out[:M])
    else:
        out[:M, :M] = ops.lin_project_gemm(Wb, M, N, Wb, M)


def jacobi_thin_svd(W, route="gram", tol=1e-13, stats=None):
    """Thin SVD of a batch of WIDE full-rank matrices W [B, M, N], M <= N, with the hand-written
    kernels only (reference: np.linalg.svd in channels/linear/linear_channel.py:8-15).

    route "gram":   G = W W^T (DMMA GEMM), rows of G rotated by block Jacobi until orthogonal:
                    they are then lambda_i u_i^T, so s_i = sqrt(lambda_i), u_i the normalised
                    rows, and V^T = diag(1/s) U^T W (DMMA GEMM).  Squares cond(W): for
                    well-conditioned W (cond^2 << 1/eps).
    route "direct": the rows of W themselves are rotated until orthogonal: they are then
                    s_i v_i^T; U^T = diag(1/s) V^T W^T (DMMA GEMM).  No squaring.
    Returns (Ut [B, M, M], s [B, M], Vt [B, M, N]), s descending."""
    t = ops.torch()
    B, M, N = W.shape
    assert M <= N
    W = W.contiguous()
    if N % 2:                                     # the GEMM kernels want 16-byte rows
        Wp = t.zeros((B, M, N + 1), dtype=t.float64, device=W.device)
        Wp[:, :, :N] = W
        W = Wp
    n_rows = _ceil_to(M, ops.JACOBI_ROWS)
    L = M if route == "gram" else N
    ld = _ceil_to(L, ops.JACOBI_COLS)
    A = t.zeros((B, n_rows, ld), dtype=t.float64, device=W.device)
    if route == "gram":
        for b in range(B):
            _gram_rows(W[b], A[b])
    else:
        A[:, :M, :N] = W[:, :, :N]
    history = jacobi_orthogonalise_rows(A, tol=tol)
    LAST_SETUP_STATS.clear()
    LAST_SETUP_STATS.update(route=route, sweeps=len(history), off=history, instances=B, rows=n_rows, ld=ld)
    if stats is not None:
        stats.update(LAST_SETUP_STATS)
    norms = ops.row_norms(A, L)
    top, order = norms.sort(dim=-1, descending=True)
    top, order = top[:, :M].contiguous(), order[:, :M].contiguous()      # padding rows have norm 0
    inv = t.where(top > 0, 1.0 / top, t.zeros_like(top))
    first = ops.rows_gather_scale(A, L, perm=order, scale=inv)           # unit rows, sorted
    del A
    if route == "gram":
        Ut, s = first, top.sqrt()
        inv_s = t.where(s > 0, 1.0 / s, t.zeros_like(s))
        Vt = t.empty((B, M, N), dtype=t.float64, device=W.device)
        for b in range(B):                                                # V^T = U^T W, then rows / s
            ops.lin_expand_gemm(W[b], M, N, Ut[b], M, out=Vt[b])
        ops.rows_gather_scale(Vt, N, scale=inv_s, out=Vt)
        return Ut, s.contiguous(), Vt
    Vt, s = first, top
    inv_s = inv
    Ut = t.empty((B, M, M), dtype=t.float64, device=W.device)
    for b in range(B):                                                    # U^T = V^T W^T, then rows / s
        ops.lin_project_gemm(W[b], M, N, Vt[b], M, out=Ut[b])
    ops.rows_gather_scale(Ut, M, scale=inv_s, out=Ut)
    return Ut, s.contiguous(), Vt


def thin_svd_device(W, method="auto"):
    """W: device tensor [B, M, N] -> (Ut [B,R,M], s [B,R], Vt [B,R,N]), s descending.

    method "jacobi" / "jacobi_direct": the hand-written set-up (`jacobi_thin_svd`, block Jacobi
    and GEMMs on the FP64 tensor cores, whole batch per launch), through the Gram matrix of
    the short side / on W itself.  Full-rank W.
    method "auto" (default): "jacobi" when W is far from square (cond(W)^2 <= 1e4: the
    singular vectors stay orthonormal to ~1e-12 and the numerical rank is unambiguous),
    "jacobi_direct" when it is nearly square or the Gram route turns out ill-conditioned, and
    the library SVD only for a rank-deficient W.
    method "svd": cuSOLVER SVD through torch.linalg (any W; the round-1 baseline).
    method "gram": torch.linalg.eigh of the smaller Gram matrix (the round-1 baseline of
    the Gram route, kept for tools/bench_setup.py).
    Setup is outside the EP sweep (the reference reports it separately as svd_time,
    examples/figures/compute_benchmark.py:27) but inside its end-to-end time
    (examples/figures/benchmark.py:22)."""
    t = ops.torch()
    B, M, N = W.shape
    if M > N and method in ("auto", "jacobi", "jacobi_direct"):
        Vt, s, Ut = thin_svd_device(W.transpose(1, 2).contiguous(), method)
        return Ut, s, Vt
    if method == "auto":
        # a nearly square matrix with independent entries has cond ~ 1 / (1 - sqrt(aspect))
        # (Marchenko-Pastur edge): beyond aspect 0.92 cond^2 exceeds 1e4
        if M <= 0.92 * N:
            Ut, s, Vt = jacobi_thin_svd(W, "gram")
            if t.isfinite(s).all() and bool(((s[:, -1] / s[:, 0])**2 >= 1e-4).all()):
                return Ut, s, Vt
        Ut, s, Vt = jacobi_thin_svd(W, "direct")
        if t.isfinite(s).all() and bool((s[:, -1] >= 1e-10 * s[:, 0]).all()):
            return Ut, s, Vt
        return thin_svd_device(W, "svd")            # rank deficient: U = W V / s is undefined
    if method in ("jacobi", "jacobi_direct"):
        # bound the work matrices of a chunk of instances to a few GB
        side = M if method == "jacobi" else N
        chunk = max(1, int(2**32 // (8 * _ceil_to(M, 32) * _ceil_to(side, 64))))
        if B > chunk:
            parts = [thin_svd_device(W[b0:b0 + chunk], method) for b0 in range(0, B, chunk)]
            return tuple(t.cat([p[k] for p in parts]) for k in range(3))
        return jacobi_thin_svd(W, "gram" if method == "jacobi" else "direct")
    if method == "svd":
        U, s, Vh = t.linalg.svd(W, full_matrices=False)
        return U.transpose(1, 2).contiguous(), s.contiguous(), Vh.contiguous()
    if method != "gram":
        raise ValueError(f"unknown svd method {method!r}")
    if M <= N:
        G = W @ W.transpose(1, 2)
        ev, U = t.linalg.eigh(G)
        ev, U = ev.flip(-1), U.flip(-1)
        s = ev.clamp_min(0).sqrt()
        Ut = U.transpose(1, 2).contiguous()
        Vt = (Ut @ W) / s[:, :, None]
        return Ut, s, Vt.contiguous()
    G = W.transpose(1, 2) @ W
    ev, V = t.linalg.eigh(G)
    ev, V = ev.flip(-1), V.flip(-1)
    s = ev.clamp_min(0).sqrt()
    Vt = V.transpose(1, 2).contiguous()
    Ut = (Vt @ W.transpose(1, 2)) / s[:, :, None]
    return Ut.contiguous(), s, Vt


class LinearChannel(Channel):
    """Linear channel x = W z.

    Parameters
    ----------
    - W: array of shape (Nx, Nz), or (B, Nx, Nz) for B independent instances
      (numpy array or device tensor)
    - precompute_svd: the reference's `False` branch keeps C = W^T W and solves the dense system
      (az I + ax C) rz = bz + W^T bx at every call (:43-45, :79-82); the thin-SVD operators give the
      same rz (it IS that solve, diagonalised), so both values take the same kernels here and the
      flag only records the caller's choice.  The factorisation is lazy either way (first use).
      One deliberate difference: for Nx < Nz the reference's `False` branch takes `singular` from the
      ASCENDING eigenvalues of C (`eigvalsh`, :45-46), i.e. the rank smallest ones including the
      Nz - Nx zeros, which makes its variances inconsistent with its own means; here `singular`
      always holds the rank non-zero eigenvalues.
    - name: str, name of weight matrix W for display
    - svd_method: "auto" | "jacobi" | "jacobi_direct" | "svd" | "gram" (extension, see
      thin_svd_device; "auto" = the hand-written block-Jacobi set-up)
    """

    def __init__(self, W, precompute_svd=True, name="W", svd_method="auto", keep_W=True):
        self.name = name
        self.Nx = int(W.shape[-2])
        self.Nz = int(W.shape[-1])
        self.precompute_svd = precompute_svd
        self.repr_init()
        self.batch = int(W.shape[0]) if len(W.shape) == 3 else None
        self.group = None
        self.W = W if keep_W else None
        self.alpha = self.Nx / self.Nz
        self._ops_ready = False
        self._svd_method = svd_method
        self._W_for_setup = W

    # ---- device setup (lazy: building a model never needs the GPU) ----------
    @classmethod
    def from_factors(cls, Ut, s, Vt, Nx, Nz, rank=None, name="W"):
        """Build from thin-SVD factors already on the device:
        Ut [Bop, R, ldm], s [Bop, R], Vt [Bop, R, ldn] (rows zero-padded)."""
        self = cls.__new__(cls)
        self.name, self.Nx, self.Nz, self.precompute_svd = name, int(Nx), int(Nz), True
        self.repr_init()
        self.batch = int(Ut.shape[0]) if Ut.shape[0] > 1 else None
        self.group = None
        self.W = None
        self.alpha = self.Nx / self.Nz
        self._install(Ut, s, Vt, rank)
        return self

    @classmethod
    def from_sharded_factors(cls, Ut, s, Vt, s_full, Nx, Nz, group, rank=None, name="W"):
        """Row shard of a single large operator (BASELINE config 5): this
        rank holds the singular triplets `Ut [1, Rg, ldm], s [1, Rg], Vt [1, Rg, ldn]`
        of W and the full spectrum `s_full [R_total]` (replicated).  Projections
        and the spectrum rescale are local; the two expansions per iteration are
        partial sums that `all_reduce` adds over `group` (NCCL over NVLink)."""
        t = ops.torch()
        self = cls.from_factors(Ut, s, Vt, Nx, Nz, rank=len(s_full) if rank is None else rank, name=name)
        self.batch = None
        self.group = group
        self.s_full = ops.to_dev(s_full).reshape(1, -1).contiguous()
        self.s2_full = (self.s_full * self.s_full).contiguous()
        self.R_total = int(self.s_full.shape[1])
        # peer-memory exchange of the two expansions per iteration (trb_comm_*)
        from ..distributed import PeerExchange
        self.exchange = PeerExchange(max(self.ldn, self.ldm), group)
        return self

    def all_reduce(self, tensor):
        """Library all-reduce of a partial expansion (the "sharded_nccl" baseline)."""
        import torch.distributed as dist
        from ..distributed import collective_device
        if collective_device(self.group) == "cpu" and tensor.is_cuda:     # gloo: staged through the host
            host = tensor.cpu()
            dist.all_reduce(host, op=dist.ReduceOp.SUM, group=self.group)
            tensor.copy_(host)
            return tensor
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group)
        return tensor

    def _install(self, Ut, s, Vt, rank=None):
        t = ops.torch()
        self.R = int(s.shape[-1])
        self.ldn, self.ldm = int(Vt.shape[-1]), int(Ut.shape[-1])
        self.Ut, self.Vt = Ut, Vt
        self.s = s.contiguous()
        self.s2 = (self.s * self.s).contiguous()
        if rank is None:
            # np.linalg.matrix_rank: s > s.max() * max(M, N) * eps (reference :37)
            tol = self.s.max(dim=-1, keepdim=True).values * max(self.Nx, self.Nz) * np.finfo(float).eps
            ranks = (self.s > tol).sum(dim=-1)
            rank = int(ranks.min().item())
            if int(ranks.max().item()) != rank:
                raise NotImplementedError("instances of one batch must share the same rank")
        self.rank = int(rank)
        self._ops_ready = True

    def _setup(self):
        if self._ops_ready:
            return
        t = ops.torch()
        W = ops.to_dev(self._W_for_setup)
        if W.dim() == 2:
            W = W[None]
        Ut, s, Vt = thin_svd_device(W, self._svd_method)
        B, R = s.shape
        Vt_p = t.zeros((B, R, ops.pad_ld(self.Nz)), dtype=t.float64, device=W.device)
        Vt_p[:, :, :self.Nz] = Vt
        Ut_p = t.zeros((B, R, ops.pad_ld(self.Nx)), dtype=t.float64, device=W.device)
        Ut_p[:, :, :self.Nx] = Ut
        self._install(Ut_p, s, Vt_p)
        self._W_for_setup = None

    # ---- reference attributes ----------------------------------------------
    @property
    def spectrum(self):
        """diag(S^T S), length Nz with zeros beyond R (reference :41)."""
        self._setup()
        s2 = self.s2.cpu().numpy()
        out = np.zeros(s2.shape[:-1] + (self.Nz,))
        out[..., :self.R] = s2
        return out[0] if self.batch is None else out

    @property
    def singular(self):
        return self.spectrum[..., :self.rank]

    def _dense_W(self):
        if self.W is not None:
            return np.asarray(self.W.cpu().numpy() if ops.is_tensor(self.W) else self.W)
        self._setup()
        t = ops.torch()
        W = t.einsum("brm,br,brn->bmn", self.Ut[:, :, :self.Nx], self.s, self.Vt[:, :, :self.Nz])
        W = W.cpu().numpy()
        return W[0] if self.batch is None else W

    def sample(self, Z):
        """X = W Z (reference :48-50); host-side, used only to draw teacher data
        and to infer shapes."""
        W = self._dense_W()
        Z = np.asarray(Z)
        if W.ndim == 3:
            return np.einsum("bmn,bn->bm", W, Z)
        if Z.ndim == 2:       # shared W, batch of signals
            return Z @ W.T
        return W @ Z

    def infer_shape(self, z_shape):
        """Output shape without forming W z (sample() draws nothing from the RNG,
        so skipping it in Model.init_shapes keeps seed parity)."""
        return [tuple(z_shape[:-1]) + (self.Nx,)]

    def math(self):
        return r"$" + self.name + "$"

    def second_moment(self, tau_z):
        return tau_z * self.spectrum.sum(axis=-1) / self.Nx

    # ---- EP factor API --------------------------------------------------------
    def _args(self, az, bz, ax, bx):
        self._setup()
        az_arg = _Arg(az, bz)
        ax_arg = _Arg(ax, bx)
        if az_arg.a_elementwise or ax_arg.a_elementwise:
            raise ValueError("LinearChannel needs scalar precisions az, ax (isotropic beliefs)")
        if az_arg.n != self.Nz or ax_arg.n != self.Nx:
            raise ValueError(f"expected bz of size {self.Nz} and bx of size {self.Nx}")
        if az_arg.B != ax_arg.B:
            raise ValueError("bz and bx disagree on the batch size")
        return az_arg, ax_arg

    def _column_block(self, bz, bx):
        """reference :75-76: `bz` of shape [Nz, k] (and `bx` [Nx, k]) is a block of k right-hand sides
        for ONE channel and ONE pair of scalar precisions -- not a batch of instances.  True for a
        2-D `bz` whose FIRST axis is Nz when this channel is un-batched (a batch of instances is
        [B, Nz]; a square [Nz, Nz] block is read as columns, like the reference would)."""
        return (self.batch is None and np.ndim(bz) == 2 and np.shape(bz)[0] == self.Nz
                and np.ndim(bx) == 2 and np.shape(bx)[0] == self.Nx and np.shape(bx)[1] == np.shape(bz)[1])

    def _means(self, az, bz, ax, bx, want):
        if getattr(self, "group", None) is not None:
            raise NotImplementedError("the factor-level API of a row-sharded LinearChannel is not "
                                      "available; run it through ExpectationPropagation")
        if self._column_block(bz, bx):
            if np.ndim(az) or np.ndim(ax):
                raise ValueError("a block of columns bz [Nz, k] shares one scalar az and one scalar ax")
            tr = (lambda v: v.t().contiguous()) if ops.is_tensor(bz) else (lambda v: np.ascontiguousarray(np.asarray(v).T))
            out = self._means(az, tr(bz), ax, tr(bx), want)          # k rows through the shared operator
            for key in ("rz", "rx"):
                if key in out:
                    out[key] = tr(out[key])
            for key in ("vz", "vx"):                                  # one scalar for the block, as in the reference
                if key in out:
                    out[key] = out[key][0] if ops.is_tensor(out[key]) else float(np.asarray(out[key]).reshape(-1)[0])
            return out
        zarg, xarg = self._args(az, bz, ax, bx)
        B = zarg.B
        bz_d = ops.padded(zarg.b[:, :self.Nz], self.ldn)
        bx_d = ops.padded(xarg.b[:, :self.Nx], self.ldm)
        tz = ops.lin_project(self.Vt, self.R, self.Nz, bz_d, B)     # V.T @ bz  (:73)
        tx = ops.lin_project(self.Ut, self.R, self.Nx, bx_d, B)     # U.T @ bx  (:72)
        out = {}
        if "rz" in want or "vz" in want:
            coef, vz = ops.lin_rescale(1, B, self.R, self.Nz, self.Nx, self.rank, self.s, self.s2,
                                       zarg.a, xarg.a, tz, tx)
            null = self.R < self.Nz
            rz = ops.lin_expand(self.Vt, self.R, self.Nz, coef, B,
                                add=bz_d if null else None, add_div=zarg.a if null else None)
            out["rz"], out["vz"] = zarg.vec_out(rz), zarg.scalar_out(vz)
        if "rx" in want or "vx" in want:
            coef, vx = ops.lin_rescale(0, B, self.R, self.Nz, self.Nx, self.rank, self.s, self.s2,
                                       zarg.a, xarg.a, tz, tx)
            rx = ops.lin_expand(self.Ut, self.R, self.Nx, coef, B)
            out["rx"], out["vx"] = xarg.vec_out(rx), xarg.scalar_out(vx)
        out["_tz"], out["_tx"], out["_zarg"], out["_xarg"] = tz, tx, zarg, xarg
        return out

    def compute_backward_mean(self, az, bz, ax, bx):
        """reference :69-83."""
        return self._means(az, bz, ax, bx, ("rz",))["rz"]

    def compute_forward_mean(self, az, bz, ax, bx):
        """reference :85-89."""
        return self._means(az, bz, ax, bx, ("rx",))["rx"]

    def _variance(self, direction, az, ax):
        self._setup()
        t = ops.torch()
        az_d, ax_d = ops.to_dev(az).reshape(-1), ops.to_dev(ax).reshape(-1)
        B = max(az_d.numel(), ax_d.numel())
        az_d, ax_d = az_d.expand(B).contiguous(), ax_d.expand(B).contiguous()
        if self.s.shape[0] not in (1, B):
            raise ValueError("az/ax batch does not match the operator batch")
        zero = t.zeros((B, self.R), dtype=t.float64, device=self.s.device)
        _, v = ops.lin_rescale(direction, B, self.R, self.Nz, self.Nx, self.rank, self.s, self.s2,
                               az_d, ax_d, zero, zero)
        numpy_out = not (ops.is_tensor(az) or ops.is_tensor(ax))
        if numpy_out:
            v = v.cpu().numpy()
            return float(v[0]) if (np.ndim(az) == 0 and np.ndim(ax) == 0) else v
        return v

    def compute_backward_variance(self, az, ax):
        """reference :91-97."""
        return self._variance(1, az, ax)

    def compute_forward_variance(self, az, ax):
        """reference :99-105."""
        return self._variance(0, az, ax)

    def compute_n_eff(self, az, ax):
        """reference :58-67, recovered from the backward variance identity is
        avoided: evaluated directly on the host copy of the spectrum (cold path)."""
        if ax == 0:
            return 0.
        if az / ax == 0:
            return self.rank / self.Nz
        singular = self.singular
        return np.sum(singular / (az / ax + singular), axis=-1) / self.Nz

    def compute_backward_posterior(self, az, bz, ax, bx):
        """reference :107-111."""
        o = self._means(az, bz, ax, bx, ("rz", "vz"))
        return o["rz"], o["vz"]

    def compute_forward_posterior(self, az, bz, ax, bx):
        """reference :113-117."""
        o = self._means(az, bz, ax, bx, ("rx", "vx"))
        return o["rx"], o["vx"]

    def compute_log_partition(self, az, bz, ax, bx):
        """reference :127-132 (a SUM): 0.5 sum(b rz) + 0.5 sum log(2 pi / a) with
        b = bz + W^T bx, a = az + ax spectrum.  Cold path (log_evidence): the
        GEMVs run in the CUDA kernels, the final dot/log-sum in torch."""
        t = ops.torch()
        o = self._means(az, bz, ax, bx, ("rz",))
        zarg, xarg = o["_zarg"], o["_xarg"]
        B = zarg.B
        rz = ops.to_dev(o["rz"]).reshape(B, self.Nz)
        wt_bx = ops.lin_expand(self.Vt, self.R, self.Nz,
                               (self.s * o["_tx"]).contiguous(), B)[:, :self.Nz]
        b = zarg.b[:, :self.Nz] + wt_bx
        a = zarg.a[:, None] + xarg.a[:, None] * self.s2
        logZ = 0.5 * (b * rz).sum(-1) + 0.5 * t.log(2 * np.pi / a).sum(-1) \
            + 0.5 * (self.Nz - self.R) * t.log(2 * np.pi / zarg.a)
        return zarg.scalar_out(logZ)

    # ---- State Evolution (reference linear_channel.py:119-143) -----------------
    def compute_backward_error(self, az, ax, tau_z):
        return self.compute_backward_variance(az, ax)

    def compute_forward_error(self, az, ax, tau_z):
        return self.compute_forward_variance(az, ax)

    def compute_mutual_information(self, az, ax, tau_z):
        """mean over the Nz eigenvalues of 0.5 log((az + ax spectrum) tau_z) (:134-137);
        cold path, reduced on the device copy of the spectrum.  A batched channel returns one
        value per instance (az, ax, tau_z scalars or arrays [B])."""
        self._setup()
        t = ops.torch()
        B = self.s2.shape[0]
        col = lambda v: ops.to_dev(np.broadcast_to(np.asarray(v, dtype=float), (B,)).copy())   # noqa: E731
        az_d, ax_d, tz_d = col(az), col(ax), col(tau_z)
        logs = t.log((az_d[:, None] + ax_d[:, None] * self.s2) * tz_d[:, None]).sum(-1) \
            + (self.Nz - self.R) * t.log(az_d * tz_d)
        out = (0.5 * logs / self.Nz).cpu().numpy()
        return float(out[0]) if self.batch is None else out

    def compute_free_energy(self, az, ax, tau_z):
        tau_x = self.second_moment(tau_z)
        I = self.compute_mutual_information(az, ax, tau_z)
        return 0.5 * (az * tau_z + self.alpha * ax * tau_x) - I + 0.5 * np.log(2 * np.pi * tau_z / np.e)
