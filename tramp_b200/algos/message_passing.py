"""The message-schedule driver (reference tramp/algos/message_passing.py).

`MessagePassing.iterate(max_iter, callback, initializer, damping, warm_start)`
keeps the reference's signature and semantics for the chain
`prior -> x -> LinearChannel -> z -> likelihood`, but the sweep itself is
device resident: all eight edge messages, both posteriors and the per-iteration
records live in HBM and one C call (`trb_sweep_run`) enqueues every kernel of
every iteration; there is no per-factor host round trip (reference :249-269
loops over nodes in Python).

Edge numbering: e1 prior->x, e2 x->lin, e3 lin->z, e4 z->lik (fwd);
e5 lik->z, e6 z->lin, e7 lin->x, e8 x->prior (bwd).
"""
import ctypes as C
import logging
import numpy as np

from .callbacks import Callback
from .initial_conditions import ConstantInit
from ..models import Model
from ..base import Variable
from ..priors import Prior
from ..likelihoods import Likelihood
from ..channels import LinearChannel
from ..variables import SISOVariable
from .. import ops, _lib

logger = logging.getLogger(__name__)

# (edge name, variable role, direction, index into edge_a)
EDGES = [("e1", "x", "fwd", 0), ("e2", "x", "fwd", 1), ("e3", "z", "fwd", 2), ("e4", "z", "fwd", 3),
         ("e5", "z", "bwd", 4), ("e6", "z", "bwd", 5), ("e7", "x", "bwd", 6), ("e8", "x", "bwd", 7)]


class MessageSnapshot:
    """Copy of the device-resident message state (the role `message_dag.copy()`
    plays in the reference, message_passing.py:356, callbacks.py:286)."""

    def __init__(self, tensors, n_iter):
        self.tensors = tensors
        self.n_iter = n_iter


INIT_ORDER = ("e1", "e2", "e8", "e3", "e7", "e4", "e6", "e5")


class MessagePassing():

    def __init__(self, model, message_keys):
        if not isinstance(model, Model):
            raise ValueError(f"model {model} is not a Model")
        self.message_keys = message_keys
        self.model = model
        self.model_dag = model.dag
        self.forward_ordering = model.forward_ordering
        self.backward_ordering = list(reversed(model.forward_ordering))
        self.variables = model.variables
        self.n_iter = 0
        self.gemv_impl = 0
        # how the four operator passes run: "gemv" (batched HBM-bound GEMVs, the
        # default), "gemm" (hand-written FP64 tensor-core DMMA GEMMs when a batch
        # shares one W), "sharded" (rows of the operators split over ranks, the
        # expansions exchanged through peer memory inside the update kernels),
        # "cublas" (torch.matmul) and "sharded_nccl" (local GEMVs + NCCL all-reduce)
        # are kept only as the library baselines the native paths are measured
        # against; None = pick
        self.linear_backend = None
        # operator passes per iteration: "general" (4, any likelihood), "gauss3" /
        # "gauss2" (3 / 2 passes, exact for a Gaussian likelihood, see
        # include/tramp_b200.h trb_sweep.schedule), "auto" = the cheapest one the
        # model, the damping and the callback allow
        self.schedule = "auto"
        self._state = None
        self._has_messages = False
        self._host = None            # FactorSchedule while the host-driven path owns the messages
        self.update_dA = False
        self.adaptive_damping = False
        self._compile_chain()

    # ------------------------------------------------------------------ model
    def _compile_chain(self):
        order = self.forward_ordering
        ok = (len(order) == 5 and isinstance(order[0], Prior) and isinstance(order[1], SISOVariable)
              and isinstance(order[2], LinearChannel) and isinstance(order[3], SISOVariable)
              and isinstance(order[4], Likelihood))
        if not ok:
            raise NotImplementedError(
                "tramp_b200 runs EP on the generalized linear model "
                "prior @ V @ LinearChannel @ V @ likelihood (observed); got "
                + " -> ".join(type(n).__name__ for n in order))
        self.prior, self.x_var, self.linear, self.z_var, self.lik = order
        self.x_id, self.z_id = self.x_var.id, self.z_var.id
        self.variable_ids = [self.x_id, self.z_id]
        if not getattr(self.prior, "isotropic", True) or not getattr(self.lik, "isotropic", True):
            raise NotImplementedError("the EP sweep uses isotropic beliefs (one precision per edge)")
        batches = {b for b in (self.prior.batch, self.linear.batch, self.lik.batch) if b is not None}
        if len(batches) > 1:
            raise ValueError(f"inconsistent batch sizes {sorted(batches)} in the model")
        self.batched = bool(batches)
        self.B = batches.pop() if batches else 1
        if self.batched and (self.prior.batch is None or self.lik.batch is None):
            raise ValueError("a batched model needs prior(batch=B) and y of shape (B, M)")
        self.N, self.M = self.linear.Nz, self.linear.Nx
        size = self.prior.size
        if (size if isinstance(size, int) else int(np.prod(size))) != self.N:
            raise ValueError(f"prior size {size} does not match W with Nz={self.N}")
        if self.lik.y is None or np.shape(self.lik.y)[-1] != self.M:
            raise ValueError(f"likelihood y must have last dimension Nx={self.M}")

    def _var_shape(self, role):
        n = self.N if role == "x" else self.M
        return (self.B, n) if self.batched else (n,)

    # ------------------------------------------------------------ device state
    def _ensure_state(self):
        if self._state is not None:
            return self._state
        t = ops.torch()
        lin = self.linear
        lin._setup()
        if lin.s.shape[0] not in (1, self.B):
            raise ValueError("operator batch does not match the model batch")
        B, R, ldn, ldm = self.B, lin.R, lin.ldn, lin.ldm
        f64 = dict(dtype=t.float64, device=lin.s.device)
        i32 = dict(dtype=t.int32, device=lin.s.device)
        st = dict(
            edge_a=t.zeros((8, B), **f64),
            b1=t.zeros((B, ldn), **f64), b7=t.zeros((B, ldn), **f64),
            b3=t.zeros((B, ldm), **f64), b5=t.zeros((B, ldm), **f64),
            rx=t.zeros((B, ldn), **f64), rz=t.zeros((B, ldm), **f64),
            vx=t.zeros(B, **f64), vz=t.zeros(B, **f64),
            tz=t.zeros((B, R), **f64), tx=t.zeros((B, R), **f64), coef=t.zeros((B, R), **f64),
            scr_n=t.zeros((B, ldn), **f64), scr_m=t.zeros((B, ldm), **f64),
            vlin=t.zeros(B, **f64), stats=t.zeros((B, 4), **f64),
            active=t.ones(B, **i32), flags=t.zeros(B, **i32), n_iter=t.zeros(B, **i32),
        )
        for k in ("edge_a", "b1", "b3", "b5", "b7", "rx", "rz", "vx", "vz", "tx"):
            st["snap_" + k] = t.zeros_like(st[k])
        self.backend = self._pick_backend()
        st["nslots"] = ops.lin_expand_slots(B, R)
        st["part"] = t.zeros((B, st["nslots"], max(ldn, ldm)), **f64)
        if self.backend not in ("gemv", "gemm", "sharded"):
            # fully reduced expansion results, read by the update kernels as slot 0
            st["red"] = t.zeros(B * max(ldn, ldm), **f64)
            st["red_n"] = st["red"][:B * ldn].view(B, ldn)
            st["red_m"] = st["red"][:B * ldm].view(B, ldm)
        y = np.asarray(self.lik.y, dtype=np.float64) if not ops.is_tensor(self.lik.y) else self.lik.y
        y2 = y if len(y.shape) == 2 else y[None, :]
        st["y"] = ops.padded(y2, ldm)
        st["b6_init"] = None
        st["b8_init"] = None
        st["x_true"] = None
        self._state = st
        return st

    def _pick_backend(self):
        lin = self.linear
        if self.linear_backend is not None:
            return self.linear_backend
        if getattr(lin, "group", None) is not None:
            return "sharded"
        # many instances sharing one W: the passes are dense [B, n] x [n, R] products
        if lin.s.shape[0] == 1 and self.B >= 16:
            return "gemm"
        return "gemv"

    def _vec_to_dev(self, value, role):
        n, ld = (self.N, self.linear.ldn) if role == "x" else (self.M, self.linear.ldm)
        v = ops.to_dev(value)
        if v.dim() == 0:
            v = v.expand(self.B, n)
        elif v.dim() == 1:
            v = v[None, :].expand(self.B, n)
        return ops.padded(v.contiguous(), ld)

    def init_message_dag(self, initializer):
        """reference message_passing.py:211-232: every edge gets a, b from the initializer."""
        st = self._ensure_state()
        t = ops.torch()
        ids = {"x": self.x_id, "z": self.z_id}
        if type(initializer) is ConstantInit and np.ndim(initializer.a) == 0 \
                and np.ndim(initializer.b) == 0:
            # same messages as the generic path below, without shipping
            # B x N constants over PCIe
            st["edge_a"].fill_(float(initializer.a))
            for k in ("b1", "b3", "b5", "b7"):
                st[k].fill_(float(initializer.b))
            st["b6_init"] = st["b8_init"] = None
            st["b6_zero"] = float(initializer.b) == 0.0
            for k in ("rx", "rz", "vx", "vz"):
                st[k].zero_()
            self._has_messages = True
            return
        init_b = {}
        # The reference walks `message_dag.edges()` (message_passing.py:223-230), i.e. node by node
        # in the order the nodes were added, out-edges in the order they were added (forward edges,
        # then backward ones): e1, e2, e8, e3, e7, e4, e6, e5.  A NoisyInit seeded like the
        # reference's therefore draws the same initial messages.
        for name, role, direction, idx in sorted(EDGES, key=lambda e: INIT_ORDER.index(e[0])):
            shape = self._var_shape(role)
            a = initializer.init("a", shape, ids[role], direction)
            b = initializer.init("b", shape, ids[role], direction)
            st["edge_a"][idx] = ops.to_dev(np.broadcast_to(np.asarray(a, dtype=np.float64), (self.B,)).copy())
            init_b[name] = self._vec_to_dev(b, role)
        st["b1"].copy_(init_b["e1"])
        st["b3"].copy_(init_b["e3"])
        st["b5"].copy_(init_b["e5"])
        st["b7"].copy_(init_b["e7"])
        # e6 / e8 are read once (first F3 / F1) before the pass-through overwrites
        # them; keep them separately only if they differ from e5 / e7
        st["b6_zero"] = not bool(init_b["e6"].any().item())
        st["b6_init"] = None if t.equal(init_b["e6"], init_b["e5"]) else init_b["e6"]
        st["b8_init"] = None if t.equal(init_b["e8"], init_b["e7"]) else init_b["e8"]
        for k in ("rx", "rz", "vx", "vz"):
            st[k].zero_()
        self._has_messages = True

    def configure_damping(self, damping):
        """reference message_passing.py:70-106: None | float | list of
        (variable.id, direction, damping) for the factor->variable edges."""
        self.damp = dict(e1=0.0, e3=0.0, e5=0.0, e7=0.0)
        self.adaptive_damping = False
        if not damping:
            self.damping = False
            return
        self.damping = True
        if isinstance(damping, str) and damping == "adaptive":
            # :151-185 -- needs the objective after every message: host-driven
            # factor-by-factor schedule (algos/factor_schedule.py)
            self.adaptive_damping = True
            return
        if not (isinstance(damping, float) or isinstance(damping, list)):
            raise ValueError("damping must be 'adaptive', float or list")
        if isinstance(damping, float):
            damping = [(x_id, d, damping) for d in ("fwd", "bwd") for x_id in self.variable_ids]
        into = {(self.x_id, "fwd"): "e1", (self.x_id, "bwd"): "e7",
                (self.z_id, "fwd"): "e3", (self.z_id, "bwd"): "e5"}
        for id, direction, damp in damping:
            if (id, direction) not in into:
                raise ValueError(f"no factor->variable edge into {id!r} with direction {direction!r}")
            self.damp[into[(id, direction)]] = float(damp or 0.0)

    def _descriptor(self, rec=None, max_records=0, early=None, synchronous=False):
        st = self._ensure_state()
        lin = self.linear
        p = _lib.ptr
        sw = _lib.TrbSweep()
        sw.B, sw.N, sw.M, sw.R = self.B, self.N, self.M, lin.R
        sw.ldn, sw.ldm, sw.rank, sw.nslots = lin.ldn, lin.ldm, lin.rank, st["nslots"]
        sw.prior = self.prior._trb_factor()
        sw.lik = self.lik._trb_factor()
        sw.lin_amin, sw.lin_amax = lin.AMIN, lin.AMAX
        shared = lin.s.shape[0] == 1
        sw.Vt, sw.strideV = p(lin.Vt), 0 if shared else lin.Vt.stride(0)
        sw.Ut, sw.strideU = p(lin.Ut), 0 if shared else lin.Ut.stride(0)
        sw.s, sw.s2, sw.stride_s = p(lin.s), p(lin.s2), 0 if shared else lin.s.stride(0)
        sw.y, sw.x_true = p(st["y"]), p(st["x_true"])
        sw.edge_a = p(st["edge_a"])
        sw.b1, sw.b3, sw.b5, sw.b7 = p(st["b1"]), p(st["b3"]), p(st["b5"]), p(st["b7"])
        sw.b6_init, sw.b8_init = p(st["b6_init"]), p(st["b8_init"])
        sw.damp1, sw.damp3, sw.damp5, sw.damp7 = (self.damp[k] for k in ("e1", "e3", "e5", "e7"))
        sw.rx, sw.rz, sw.vx, sw.vz = p(st["rx"]), p(st["rz"]), p(st["vx"]), p(st["vz"])
        sw.tz, sw.tx, sw.coef, sw.part = p(st["tz"]), p(st["tx"]), p(st["coef"]), p(st["part"])
        sw.scr_n, sw.scr_m, sw.vlin, sw.stats = p(st["scr_n"]), p(st["scr_m"]), p(st["vlin"]), p(st["stats"])
        sw.active, sw.flags, sw.n_iter = p(st["active"]), p(st["flags"]), p(st["n_iter"])
        rec = rec or {}
        sw.rec_mse, sw.rec_smse = p(rec.get("mse")), p(rec.get("smse"))
        sw.rec_vx, sw.rec_vz, sw.rec_tol = p(rec.get("vx")), p(rec.get("vz")), p(rec.get("tol"))
        sw.max_records = max_records
        if early is not None:
            sw.es_tol, sw.es_max_increase = early.tol, early.max_increase
            sw.es_wait_increase, sw.es_vars = early.wait_increase, early._var_mask(self)
            # EarlyStopping (variances) or EarlyStoppingEP (means): trb_sweep.es_mode
            sw.es_mode = 1 if hasattr(early, "min_variance") else 0
            sw.es_min_variance = getattr(early, "min_variance", -1.0)
        else:
            sw.es_tol, sw.es_max_increase, sw.es_wait_increase, sw.es_vars = -1.0, 0.0, 0, 3
        sw.gemv_impl = 3 if self.backend == "gemm" else (self.gemv_impl or 2)
        sw.R_total = getattr(lin, "R_total", 0) or 0
        if self.backend == "sharded":
            sw.comm = lin.exchange.ptr
            sw.s_full, sw.s2_full = p(lin.s_full), p(lin.s2_full)
        for k in ("edge_a", "b1", "b3", "b5", "b7", "rx", "rz", "vx", "vz", "tx"):
            setattr(sw, "snap_" + k, p(st["snap_" + k]))
        sw.schedule = self.last_schedule = self._pick_schedule(early, synchronous)
        if sw.schedule:
            if "ty" not in st:     # ty = U_R^T y, once per model (y is fixed)
                st["ty"] = ops.torch().zeros_like(st["tx"])
                sw.ty = p(st["ty"])
                _lib.check(_lib.load().trb_sweep_stage(C.byref(sw), _lib.STAGE_PROJECT_Y, 0, 0, 0,
                                                       _lib.current_stream()))
            sw.ty = p(st["ty"])
            if sw.schedule == 2:
                sw.es_vars = 1     # the recorded tolerance covers x only
        return sw

    def _pick_schedule(self, early, synchronous=False):
        """0 general / 1 gauss3 / 2 gauss2 (trb_sweep.schedule)."""
        from ..likelihoods import GaussianLikelihood
        want = self.schedule
        if want not in ("auto", "general", "gauss3", "gauss2"):
            raise ValueError(f"unknown schedule {want!r}")
        st = self._state
        ok3 = (type(self.lik) is GaussianLikelihood and self.backend in ("gemv", "gemm", "sharded")
               and st.get("b6_init") is None)
        ok2 = ok3 and self.damp["e3"] == 0.0 and early is None and not synchronous
        if want == "general":
            return 0
        if want == "gauss3":
            if not ok3:
                raise ValueError("schedule 'gauss3' needs a Gaussian likelihood and e6 initialised like e5")
            return 1
        if want == "gauss2":
            if not ok2:
                raise ValueError("schedule 'gauss2' needs a Gaussian likelihood, no damping of the "
                                 "linear->z edge, no early stopping and a replayable callback")
            return 2
        return 2 if ok2 else (1 if ok3 else 0)

    def _run(self, sw, it0, n_iter, fresh):
        """Enqueue n_iter iterations.  fresh: the messages were just initialised."""
        st = self._state
        code = 0 if not fresh else (2 if st.get("b6_zero") else 1)
        if getattr(self, "_tx_stale", False):
            code, self._tx_stale = 1, False
        if self.backend in ("gemv", "gemm", "sharded"):
            _lib.check(_lib.load().trb_sweep_run(C.byref(sw), it0, n_iter, code, _lib.current_stream()))
        else:
            self._run_staged(sw, it0, n_iter, code)

    def _run_staged(self, sw, it0, n_iter, fresh):
        """Same schedule as trb_sweep_run (tramp_b200/csrc/trb_sweep.cu) with the four
        operator passes replaced by the back end: cuBLAS FP64 GEMMs on a shared W
        ("cublas") or local GEMVs on this rank's row shard followed by an all-reduce
        ("sharded").  Every other stage is the same CUDA kernel."""
        t = ops.torch()
        lib = _lib.load()
        st, lin = self._state, self.linear
        stream = _lib.current_stream()
        B, N, M, R = self.B, self.N, self.M, lin.R
        ea = st["edge_a"]
        # descriptor whose `part` is the fully reduced buffer (slot 0, nslots = 1)
        red = _lib.TrbSweep.from_buffer_copy(sw)
        red.part, red.nslots = _lib.ptr(st["red"]), 1
        sharded = self.backend == "sharded_nccl"

        def stage(desc, which, it, first, pre=0):
            _lib.check(lib.trb_sweep_stage(C.byref(desc), which, it, int(first), pre, stream))

        def project(A, vec, out, n):
            if sharded:
                ops.lin_project(A, R, n, vec, B, self.gemv_impl, st["active"], out=out)
            else:
                t.matmul(vec, A[0].transpose(0, 1), out=out)

        def expand(A, n, out):
            """out[B, ld] = coef @ A (+ all-reduce over the row shards)."""
            if sharded:
                ops.lin_expand(A, R, n, st["coef"], B, self.gemv_impl, st["active"], out=out,
                               part=st["part"][:, :, :A.shape[-1]] if st["part"].shape[-1] == A.shape[-1]
                               else None)
                lin.all_reduce(out)
            else:
                t.matmul(st["coef"], A[0], out=out)

        def rescale(direction):
            if sharded:
                # variance from the full spectrum (replicated), coefficients from the shard
                ops.lin_rescale(direction, B, lin.R_total, N, M, lin.rank, lin.s_full, lin.s2_full,
                                ea[1], ea[5], None, None, st["active"], null_space=lin.R_total < N,
                                want_coef=False, v=st["vlin"])
                ops.lin_rescale(direction, B, R, N, M, min(lin.rank, R), lin.s, lin.s2, ea[1], ea[5],
                                st["tz"], st["tx"], st["active"], null_space=lin.R_total < N,
                                want_v=False, coef=st["coef"])
            else:
                stage(sw, _lib.STAGE_RESCALE_FWD if direction == 0 else _lib.STAGE_RESCALE_BWD, 0, 0)

        for k in range(n_iter):
            it, first = it0 + k, bool(fresh) and k == 0
            stage(sw, _lib.STAGE_PRIOR, it, first)
            project(lin.Vt, st["b1"], st["tz"], N)                          # P1
            if first:
                if fresh == 2:
                    st["tx"].zero_()
                else:
                    project(lin.Ut, st["b6_init"] if st["b6_init"] is not None else st["b5"],
                            st["tx"], M)
            rescale(0)                                                      # S1
            expand(lin.Ut, M, st["red_m"])                                  # P2
            stage(red, _lib.STAGE_Z_UPDATE, it, first, 1)
            project(lin.Ut, st["b5"], st["tx"], M)                          # P3
            rescale(1)                                                      # S2
            expand(lin.Vt, N, st["red_n"])                                  # P4
            stage(red, _lib.STAGE_X_UPDATE, it, first, 1)
            stage(sw, _lib.STAGE_SNAPSHOT, it, first)

    def _raise_on_nan(self, flags):
        """reference message_passing.py:187-209 (check_message)."""
        if (flags & _lib.FLAG_COMM_TIMEOUT).any():
            raise _lib.TrbError("a rank of the row-sharded operator did not publish its partial sums "
                                "within the time-out; the ranks are out of step")
        bad = np.nonzero(flags & (_lib.FLAG_NAN_A | _lib.FLAG_NAN_B))[0]
        if bad.size:
            what = "a" if (flags[bad[0]] & _lib.FLAG_NAN_A) else "b"
            where = f" in instance(s) {bad.tolist()}" if self.batched else ""
            raise ValueError(f"EP message {what} is nan{where}")
        if (flags & _lib.FLAG_NEG_A).any():
            logger.warning("negative a in an EP message")

    # ---------------------------------------------------------------- iterate
    def iterate(self, max_iter=200, callback=None, initializer=None, damping=None,
                warm_start=False, update_dA=False):
        """reference message_passing.py:330-357."""
        initializer = initializer or ConstantInit(a=0, b=0)
        callback = callback or self.default_stopping
        self.update_dA = bool(update_dA)
        host_prev = self._host
        if warm_start:
            if not self._has_messages:
                raise ValueError("message dag was never initialized")
            logger.info(f"warm start with n_iter={self.n_iter} no initialization")
        else:
            logger.info(f"init message dag with {initializer}")
            self.init_message_dag(initializer)
            self.n_iter = 0
        self.configure_damping(damping)
        if getattr(self.linear, "group", None) is not None:
            # the ranks of a row-sharded operator enter the sweep together: the peer exchange
            # (trb_comm.cu) gives a silent peer one second before it flags a time-out, and host
            # skew before the first exchange (module load, allocation) must not count against it
            import torch.distributed as dist
            dist.barrier(group=self.linear.group)
        self.n_iter_per_instance = None
        st = self._ensure_state()
        st["active"].fill_(1)
        st["flags"].zero_()
        st["n_iter"].zero_()
        fresh = not warm_start
        if self.adaptive_damping or self.update_dA:
            self._iterate_host(max_iter, callback, host_prev if warm_start else None)
            logger.info(f"terminated after n_iter={self.n_iter} iterations")
            return
        if warm_start and host_prev is not None:
            self._leave_host_path(host_prev)
        self._host = None
        if isinstance(callback, Callback) and callback.device_replayable(self):
            self._iterate_device(max_iter, callback, fresh)
        else:
            self._iterate_synchronous(max_iter, callback, fresh)
        logger.info(f"terminated after n_iter={self.n_iter} iterations")

    def _iterate_device(self, max_iter, callback, fresh):
        """Whole sweep on the device; the callbacks are fed the recorded trajectory."""
        t = ops.torch()
        st = self._state
        cfg = {}
        callback.device_config(cfg)
        st["x_true"] = None
        if cfg.get("x_true") is not None:
            st["x_true"] = self._vec_to_dev(cfg["x_true"], "x")
        early = cfg.get("early_stopping")
        rec = {k: t.full((max(max_iter, 1), self.B), float("nan"), dtype=t.float64,
                         device=st["vx"].device)
               for k in ("mse", "smse", "vx", "vz", "tol")}
        sw = self._descriptor(rec, max_iter, early)
        chunk = max_iter if early is None else 16
        it = 0
        while it < max_iter:
            k = min(chunk, max_iter - it)
            self._run(sw, it, k, fresh and it == 0)
            it += k
            if early is not None and not bool(st["active"].any().item()):
                break
        n_iter = st["n_iter"].cpu().numpy()
        flags = st["flags"].cpu().numpy()
        self.flags = flags
        self.n_iter_per_instance = self.n_iter + n_iter
        n_done = int(n_iter.max()) if n_iter.size else 0
        self.n_iter += n_done
        self._raise_on_nan(flags)
        if (flags & _lib.FLAG_DIVERGED).any():
            logger.info("EarlyStoppingEP: increase above max_increase in instance(s) "
                        f"{np.nonzero(flags & _lib.FLAG_DIVERGED)[0].tolist()}; old messages restored")
        rec_h = {k: v[:n_done].cpu().numpy() for k, v in rec.items()}
        self.records = rec_h
        for i in range(n_done):
            callback.replay(self, i, max_iter, rec_h)

    def _iterate_synchronous(self, max_iter, callback, fresh):
        """Arbitrary user callback: one device sweep iteration, then the callback
        (which may read any state through get_variables_data)."""
        st = self._state
        st["x_true"] = None
        sw = self._descriptor(synchronous=True)
        try:
            for i in range(max_iter):
                self._run(sw, i, 1, fresh and i == 0)
                self._raise_on_nan(st["flags"].cpu().numpy())
                self.n_iter += 1
                stop = callback(self, i, max_iter)
                if stop:
                    return
        finally:      # also on a callback stop or an exception: never a stale count from an earlier call
            self.n_iter_per_instance = np.full(self.B, self.n_iter)

    # ------------------------------------------- host-driven factor-by-factor path
    def _iterate_host(self, max_iter, callback, previous):
        """damping="adaptive" / update_dA (reference :129-185): the reference's node-by-node loop,
        enqueued from the host.  `schedule_backend = "device"` (default): all eight messages stay on
        the device and every step is a kernel launch (algos/device_schedule.py); "host": the same
        schedule through the numpy factor API (algos/factor_schedule.py), the cross-check."""
        if getattr(self.linear, "group", None) is not None:
            raise NotImplementedError("adaptive damping / update_dA on a row-sharded operator")
        backend = getattr(self, "schedule_backend", "device")
        if backend not in ("device", "host"):
            raise ValueError(f"unknown schedule_backend {backend!r}")
        st = self._state
        if previous is not None:
            host = previous
            for name in ("e1", "e2", "e3", "e4", "e5", "e6", "e7", "e8"):   # a new iterate() may change the constant damping
                if hasattr(host, "set_damping"):
                    host.set_damping(name, self.damp.get(name) or None)
                else:
                    host.edges[name]["damping"] = self.damp.get(name) or None
        else:
            edges = {}
            src = {"e1": "b1", "e2": "b1", "e3": "b3", "e4": "b3", "e5": "b5", "e6": "b6_init",
                   "e7": "b7", "e8": "b8_init"}
            a_all = st["edge_a"].cpu().numpy() if backend == "host" else None
            for name, role, direction, idx in EDGES:
                t = st.get(src[name])
                if t is None:
                    t = st["b5" if name == "e6" else "b7"]
                n = self.N if role == "x" else self.M
                if backend == "device":
                    edges[name] = dict(a=st["edge_a"][idx], b=t, n_iter=0, damping=self.damp.get(name) or None)
                    continue
                b_host = t[:, :n].cpu().numpy().copy()
                edges[name] = dict(a=a_all[idx].copy() if self.batched else float(a_all[idx, 0]),
                                   b=b_host if self.batched else b_host[0],
                                   direction=direction, n_iter=0,
                                   damping=self.damp.get(name) or None)
            if backend == "device":
                from .device_schedule import DeviceSchedule
                host = DeviceSchedule(self, edges)
            else:
                from .factor_schedule import FactorSchedule
                host = FactorSchedule(self, edges)
        self._host = host
        try:
            for i in range(max_iter):
                host.sweep()
                self._host_to_device(host)
                self.n_iter += 1
                stop = callback(self, i, max_iter)
                if stop:
                    return
                host.old = host.copy_state()
        finally:
            self.n_iter_per_instance = np.full(self.B, self.n_iter)

    def _host_to_device(self, host):
        """Mirror the host messages and posteriors into the device state, so that
        get_variables_data / snapshots / a later device-path warm start see them."""
        st = self._state
        t = ops.torch()
        if hasattr(host, "mirror_into"):            # the device schedule: device-to-device copies
            host.mirror_into(st)
            return
        e = host.edges
        st["edge_a"].copy_(t.as_tensor(np.stack([np.broadcast_to(np.asarray(e[n]["a"], dtype=np.float64), (self.B,))
                                                  for n, _, _, _ in EDGES]), dtype=t.float64))
        for buf, name, n in (("b1", "e1", self.N), ("b3", "e3", self.M), ("b5", "e5", self.M),
                             ("b7", "e7", self.N)):
            st[buf][:, :n] = ops.to_dev(np.asarray(e[name]["b"], dtype=np.float64).reshape(self.B, n))
        for role, n, rk, vk in (("x", self.N, "rx", "vx"), ("z", self.M, "rz", "vz")):
            d = host.variables[role]
            if d:
                st[rk][:, :n] = ops.to_dev(np.asarray(d["r"], dtype=np.float64).reshape(self.B, n))
                st[vk].copy_(ops.to_dev(np.broadcast_to(np.asarray(d["v"], dtype=np.float64), (self.B,)).copy()))

    def _leave_host_path(self, host):
        """Warm start of the device sweep from messages the host path produced: the
        pass-through edges must again be copies of their sources (they always are
        unless adaptive damping held one of them back)."""
        if hasattr(host, "aliases_hold"):
            if not host.aliases_hold():
                raise NotImplementedError("device-path warm start needs e2 == e1, e4 == e3, e6 == e5, e8 == e7; "
                                          "adaptive damping left them different")
        else:
            e = host.edges
            for cp, srcn in (("e2", "e1"), ("e4", "e3"), ("e6", "e5"), ("e8", "e7")):
                if np.any(e[cp]["a"] != e[srcn]["a"]) or not np.array_equal(e[cp]["b"], e[srcn]["b"]):
                    raise NotImplementedError(
                        f"device-path warm start needs {cp} == {srcn}; adaptive damping left them different")
        self._host_to_device(host)
        st = self._state
        st["b6_init"] = st["b8_init"] = None
        st["b6_zero"] = False
        self._tx_stale = True     # tx = U_R^T b6 is recomputed by the first device iteration

    # ------------------------------------------------------------ inspection
    def snapshot(self):
        st = self._ensure_state()
        keys = ("edge_a", "b1", "b3", "b5", "b7", "rx", "rz", "vx", "vz", "tx", "tz")
        snap = MessageSnapshot({k: st[k].clone() for k in keys}, self.n_iter)
        snap.host = self._host.copy_state() if self._host is not None else None
        return snap

    def reset_message_dag(self, snapshot):
        """reference message_passing.py:234-239."""
        st = self._ensure_state()
        for k, v in snapshot.tensors.items():
            st[k].copy_(v)
        if self._host is not None and getattr(snapshot, "host", None) is not None:
            self._host.restore_state(snapshot.host)

    def _out(self, t, n=None):
        x = t.cpu().numpy()
        if n is not None:
            x = x[:, :n]
        if not self.batched:
            x = x[0]
            return float(x) if n is None else x
        return x

    def get_variables_data(self, ids="all"):
        """reference message_passing.py:271-276: {id: dict(shape, r, v)}."""
        st = self._ensure_state()
        data = {}
        if ids == "all" or self.x_id in ids:
            data[self.x_id] = dict(shape=self._var_shape("x"), r=self._out(st["rx"], self.N),
                                   v=self._out(st["vx"]))
        if ids == "all" or self.z_id in ids:
            data[self.z_id] = dict(shape=self._var_shape("z"), r=self._out(st["rz"], self.M),
                                   v=self._out(st["vz"]))
        return data

    def get_variable_data(self, id):
        data = self.get_variables_data([id])
        if id not in data:
            raise ValueError(f"id={id} not in variables")
        return data[id]

    def _edge(self, name):
        """(a, b) of one edge as host arrays."""
        if self._host is not None:
            d = self._host.edges[name]
            return d["a"], d["b"]
        st = self._ensure_state()
        _, role, direction, idx = next(e for e in EDGES if e[0] == name)
        src = {"e1": "b1", "e2": "b1", "e3": "b3", "e4": "b3",
               "e5": "b5", "e6": "b5", "e7": "b7", "e8": "b7"}[name]
        n = self.N if role == "x" else self.M
        a = st["edge_a"][idx].cpu().numpy()
        return (a if self.batched else float(a[0])), self._out(st[src], n)

    def get_nodes_data(self, keys):
        """reference message_passing.py:289-298 (after update_objective: key "A")."""
        records = []
        A = getattr(self, "A_nodes", {})
        for node in self.forward_ordering:
            is_var = isinstance(node, Variable)
            full = dict(A=A.get(node.id))
            if is_var:
                full.update(self.get_variable_data(node.id))
            record = dict(id=node.id, type="variable" if is_var else "factor")
            for key in keys:
                record[key] = full.get(key)
            record["n_iter"] = self.n_iter
            records.append(record)
        return records

    def get_edges_data(self, keys):
        """reference message_passing.py:278-287."""
        factor_of = {"e1": self.prior, "e8": self.prior, "e2": self.linear, "e7": self.linear,
                     "e3": self.linear, "e6": self.linear, "e4": self.lik, "e5": self.lik}
        into = {"e1": "e1", "e3": "e3", "e5": "e5", "e7": "e7"}
        records = []
        for name, role, direction, idx in EDGES:
            a, b = self._edge(name)
            x_id = self.x_id if role == "x" else self.z_id
            full = dict(a=a, b=b, direction=direction, n_iter=self.n_iter, tau=None,
                        shape=self._var_shape(role),
                        damping=(self.damp.get(into.get(name)) or None) if hasattr(self, "damp") else None)
            if self._host is not None:       # per-edge n_iter, A, dA, beta of the host path
                full.update({k: v for k, v in self._host.edges[name].items() if k not in ("a", "b")})
            elif hasattr(self, "A_edge_by_name"):
                full["A"] = self.A_edge_by_name.get(name)
            record = dict(x_id=x_id, f_id=factor_of[name].id)
            for key in keys:
                record[key] = full.get(key)
            records.append(record)
        return records
