"""State Evolution (reference tramp/algos/state_evolution.py:5-27).

SE is the scalar twin of EP: the same schedule (message_passing.py:249-269,
330-357) on the precisions `a` alone, each factor averaging its posterior
variance over the law of its incoming beliefs.  The reference walks the DAG node
by node in Python and evaluates every average with scipy.integrate.quad /
dblquad.  Here the whole recursion of the chain

    prior -> x -> linear channel -> z -> likelihood

runs inside one kernel launch (`trb_se_run`, tramp_b200/csrc/trb_se.cu): one CTA
per problem, fixed-node Gauss-Legendre quadrature over the device moment
routines, damping, NaN check and EarlyStopping included.  `StateEvolution`
accepts one model (the reference API) or a list of models (an extension: a whole
grid of alpha / rho / noise values in the same launch; every per-variable value
then becomes an array of length G).

The linear channel is either a `MarchenkoPasturChannel` (closed-form transforms)
or a `LinearChannel` (its own spectrum, as `TeacherStudentScenario.run_se` uses).

Edge numbering: e1 prior->x, e2 x->lin, e3 lin->z, e4 z->lik (fwd);
e5 lik->z, e6 z->lin, e7 lin->x, e8 x->prior (bwd).
"""
import ctypes as C
import logging
import numpy as np

from .callbacks import Callback, EarlyStopping
from .initial_conditions import ConstantInit
from ..models import Model
from ..priors import Prior
from ..likelihoods import Likelihood
from ..channels import LinearChannel, AnalyticalLinearChannel
from ..variables import SISOVariable
from .. import ops, _lib

logger = logging.getLogger(__name__)

EDGES = [("e1", "x", "fwd"), ("e2", "x", "fwd"), ("e3", "z", "fwd"), ("e4", "z", "fwd"),
         ("e5", "z", "bwd"), ("e6", "z", "bwd"), ("e7", "x", "bwd"), ("e8", "x", "bwd")]
EDGE_INDEX = {name: i for i, (name, _, _) in enumerate(EDGES)}
# (source role, target role) of every edge, for get_edges_data
EDGE_ENDS = {"e1": ("prior", "x"), "e2": ("x", "lin"), "e3": ("lin", "z"), "e4": ("z", "lik"),
             "e5": ("lik", "z"), "e6": ("z", "lin"), "e7": ("lin", "x"), "e8": ("x", "prior")}


class SESnapshot:
    """Copy of the scalar message state (`message_dag.copy()` in the reference)."""

    def __init__(self, edge_a, vx, vz):
        self.edge_a, self.vx, self.vz = edge_a, vx, vz


class StateEvolution():

    def __init__(self, model):
        models = list(model) if isinstance(model, (list, tuple)) else [model]
        self.batched = isinstance(model, (list, tuple))
        if not models:
            raise ValueError("no model")
        for m in models:
            if not isinstance(m, Model):
                raise ValueError(f"model {m} is not a Model")
            m.init_second_moments()                       # reference :7
        self.models = models
        self.model = models[0]
        self.message_keys = ["a"]
        self.G = len(models)
        self.n_iter = 0
        self.default_stopping = EarlyStopping()
        self.quadrature = None        # ((P, Q, kappa), (P2, Q2, kappa2)); None = ops defaults
        self._state = None
        self._has_messages = False
        self._compile_chain()

    # ------------------------------------------------------------------ model
    def _compile_chain(self):
        chains = []
        for m in self.models:
            order = m.forward_ordering
            ok = (len(order) == 5 and isinstance(order[0], Prior) and isinstance(order[1], SISOVariable)
                  and isinstance(order[2], (LinearChannel, AnalyticalLinearChannel))
                  and isinstance(order[3], SISOVariable) and isinstance(order[4], Likelihood))
            if not ok:
                raise NotImplementedError(
                    "tramp_b200 runs State Evolution on the generalized linear model "
                    "prior @ V @ (LinearChannel | MarchenkoPasturChannel) @ V @ likelihood; got "
                    + " -> ".join(type(n).__name__ for n in order))
            chains.append(order)
        self.prior, self.x_var, self.linear, self.z_var, self.lik = chains[0]
        self.chains = chains
        self.x_id, self.z_id = self.x_var.id, self.z_var.id
        self.variable_ids = [self.x_id, self.z_id]
        for order in chains:
            if (order[1].id, order[3].id) != (self.x_id, self.z_id):
                raise ValueError("all models of a batched State Evolution must use the same variable ids")
            if not getattr(order[0], "isotropic", True) or not getattr(order[4], "isotropic", True):
                raise NotImplementedError("State Evolution uses isotropic beliefs")
        analytical = [isinstance(order[2], AnalyticalLinearChannel) for order in chains]
        if any(analytical) and not all(analytical):
            raise ValueError("cannot mix analytical and empirical linear channels in one batch")
        self.analytical = analytical[0]
        if not self.analytical:
            if self.G != 1:
                raise NotImplementedError("a batch of State Evolutions needs analytical "
                                          "(Marchenko-Pastur) channels")
            if self.linear.batch is not None:
                raise NotImplementedError("State Evolution of a batched LinearChannel is not supported")
        tau = [m.get_second_moments() for m in self.models]
        self.tau_x = np.array([t[self.x_id] for t in tau], dtype=np.float64)
        self.tau_z = np.array([t[self.z_id] for t in tau], dtype=np.float64)

    def _out(self, arr):
        """[G] host array -> float for a single model, array for a batch."""
        arr = np.asarray(arr)
        return arr.copy() if self.batched else arr[0].item()

    # ------------------------------------------------------------ device state
    def _ensure_state(self):
        if self._state is not None:
            return self._state
        t = ops.torch()
        dev = ops.device()
        f64 = dict(dtype=t.float64, device=dev)
        i32 = dict(dtype=t.int32, device=dev)
        G = self.G
        st = dict(
            edge_a=t.zeros((8, G), **f64), vx=t.zeros(G, **f64), vz=t.zeros(G, **f64),
            active=t.ones(G, **i32), flags=t.zeros(G, **i32), n_iter=t.zeros(G, **i32),
            tau_x=ops.to_dev(self.tau_x), tau_z=ops.to_dev(self.tau_z),
            prior=ops.factors_to_dev([c[0]._trb_factor() for c in self.chains]),
            lik=ops.factors_to_dev([c[4]._trb_factor() for c in self.chains]),
        )
        if self.analytical:
            st["alpha"] = ops.to_dev(np.array([c[2].alpha for c in self.chains], dtype=np.float64))
            st["mean_spectrum"] = ops.to_dev(
                np.array([c[2].ensemble.mean_spectrum for c in self.chains], dtype=np.float64))
        else:
            self.linear._setup()
        self._state = st
        return st

    def _descriptor(self, rec=None, max_records=0, early=None):
        st = self._ensure_state()
        p = _lib.ptr
        se = _lib.TrbSe()
        se.G = self.G
        se.prior, se.lik = p(st["prior"]), p(st["lik"])
        se.tau_x, se.tau_z = p(st["tau_x"]), p(st["tau_z"])
        lin = self.linear
        if self.analytical:
            se.channel = _lib.SE_MARCHENKO_PASTUR
            se.alpha, se.mean_spectrum = p(st["alpha"]), p(st["mean_spectrum"])
        else:
            se.channel = _lib.SE_SPECTRUM
            se.s2, se.stride_s2 = p(lin.s2), 0
            se.R, se.Nz, se.Nx, se.rank = lin.R, lin.Nz, lin.Nx, lin.rank
        se.lin_amin, se.lin_amax = lin.AMIN, lin.AMAX
        se.damp1, se.damp3, se.damp5, se.damp7 = (self.damp[k] for k in ("e1", "e3", "e5", "e7"))
        se.edge_a, se.vx, se.vz = p(st["edge_a"]), p(st["vx"]), p(st["vz"])
        se.active, se.flags, se.n_iter = p(st["active"]), p(st["flags"]), p(st["n_iter"])
        rec = rec or {}
        se.rec_vx, se.rec_vz, se.max_records = p(rec.get("vx")), p(rec.get("vz")), max_records
        if early is not None:
            se.es_tol, se.es_min_variance = early.tol, early.min_variance
            se.es_max_increase, se.es_wait_increase = early.max_increase, early.wait_increase
            se.es_vars = early._var_mask(self)
        else:
            se.es_tol, se.es_vars = -1.0, 3
        q = self.quadrature
        se.quad = ops.quadrature(*q) if q else ops.quadrature()
        return se

    # --------------------------------------------------------------- messages
    def init_message_dag(self, initializer):
        """reference message_passing.py:211-232: `a` of every edge from the
        initializer, keyed by (variable id, direction)."""
        st = self._ensure_state()
        ids = {"x": self.x_id, "z": self.z_id}
        a0 = np.zeros((8, self.G))
        from .message_passing import INIT_ORDER     # the reference's edge order (same random stream)
        for name in INIT_ORDER:
            i, (_, role, direction) = next((k, e) for k, e in enumerate(EDGES) if e[0] == name)
            a0[i, :] = initializer.init("a", None, ids[role], direction)
        st["edge_a"].copy_(ops.to_dev(a0))
        st["vx"].zero_()
        st["vz"].zero_()
        self._has_messages = True

    def configure_damping(self, damping):
        """reference message_passing.py:70-106 (constant damping of the
        factor->variable edges; adaptive damping is an EP feature here)."""
        self.damp = dict(e1=0.0, e3=0.0, e5=0.0, e7=0.0)
        if not damping:
            self.damping = False
            return
        self.damping = True
        if isinstance(damping, str) and damping == "adaptive":
            raise NotImplementedError("adaptive damping is not implemented for State Evolution")
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

    def snapshot(self):
        st = self._ensure_state()
        return SESnapshot(st["edge_a"].clone(), st["vx"].clone(), st["vz"].clone())

    def reset_message_dag(self, snapshot):
        """reference message_passing.py:234-239."""
        st = self._ensure_state()
        st["edge_a"].copy_(snapshot.edge_a)
        st["vx"].copy_(snapshot.vx)
        st["vz"].copy_(snapshot.vz)

    # ---------------------------------------------------------------- iterate
    def iterate(self, max_iter=200, callback=None, initializer=None, damping=None,
                warm_start=False, update_dA=False):
        """reference message_passing.py:330-357."""
        initializer = initializer or ConstantInit(a=0, b=0)
        callback = callback or self.default_stopping
        if update_dA:
            raise NotImplementedError("update_dA is not implemented for State Evolution")
        self.configure_damping(damping)
        if warm_start:
            if not self._has_messages:
                raise ValueError("message dag was never initialized")
            logger.info(f"warm start with n_iter={self.n_iter} no initialization")
        else:
            logger.info(f"init message dag with {initializer}")
            self.init_message_dag(initializer)
            self.n_iter = 0
        st = self._ensure_state()
        st["active"].fill_(1)
        st["flags"].zero_()
        st["n_iter"].zero_()
        if isinstance(callback, Callback) and callback.device_replayable(self):
            self._iterate_device(max_iter, callback)
        else:
            self._iterate_synchronous(max_iter, callback)
        logger.info(f"terminated after n_iter={self.n_iter} iterations")

    def _raise_on_flags(self, flags, grid=False):
        domain = (flags & _lib.FLAG_SE_DOMAIN) != 0
        if domain.any():
            # sgn_likelihood.py:80-81 / abs_likelihood.py:57-58.  One run: the reference's
            # AssertionError.  A grid of runs in one launch (our addition): the other
            # problems are valid results, so the failed ones read v = NaN (and keep
            # their flag in `self.flags`) instead of discarding the whole grid.
            if not (grid and self.batched) or domain.all():
                raise AssertionError("az must be greater than 1/ tau_z")
            bad = np.nonzero(domain)[0]
            logger.warning(f"az must be greater than 1/ tau_z in problem(s) {bad.tolist()}: v = nan there")
            t = ops.torch()
            idx = t.as_tensor(bad, device=self._state["vx"].device)
            for key in ("vx", "vz"):
                self._state[key].index_fill_(0, idx, float("nan"))
        if (flags & _lib.FLAG_NAN_A).any():
            bad = np.nonzero(flags & _lib.FLAG_NAN_A)[0]
            where = f" in problem(s) {bad.tolist()}" if self.batched else ""
            raise ValueError(f"SE message a is nan{where}")     # message_passing.py:190-198
        if (flags & _lib.FLAG_NEG_A).any():
            logger.warning("negative a in an SE message")

    def _iterate_device(self, max_iter, callback):
        """All iterations in one launch; callbacks replay the recorded trajectory."""
        t = ops.torch()
        st = self._state
        cfg = {}
        callback.device_config(cfg)
        early = cfg.get("early_stopping")
        n_rec = max(max_iter, 1)
        rec = {k: t.full((n_rec, self.G), float("nan"), dtype=t.float64, device=st["vx"].device)
               for k in ("vx", "vz")}
        se = self._descriptor(rec, n_rec, early)
        _lib.check(_lib.load().trb_se_run(C.byref(se), 0, max_iter, _lib.current_stream()))
        n_iter = st["n_iter"].cpu().numpy()
        flags = st["flags"].cpu().numpy()
        self.flags = flags
        self.n_iter_per_problem = self.n_iter + n_iter
        done = int(n_iter.max()) if n_iter.size else 0
        host_rec = {k: v[:max(done, 1)].cpu().numpy() for k, v in rec.items()}
        self.records = host_rec
        first = self.n_iter
        self._raise_on_flags(flags, grid=True)
        for i in range(done):
            self.n_iter = first + i + 1
            callback.replay(self, i, max_iter, host_rec)
        self.n_iter = first + done

    def _iterate_synchronous(self, max_iter, callback):
        """Any other callback: one launch per iteration, the callback sees the
        state after each (reference message_passing.py:345-356)."""
        st = self._state
        se = self._descriptor()
        lib = _lib.load()
        for i in range(max_iter):
            st["active"].fill_(1)
            _lib.check(lib.trb_se_run(C.byref(se), 0, 1, _lib.current_stream()))
            flags = st["flags"].cpu().numpy()
            self.flags = flags
            self._raise_on_flags(flags)
            self.n_iter += 1
            if callback(self, i, max_iter):
                logger.info(f"terminated after n_iter={self.n_iter} iterations")
                return

    # ------------------------------------------------------------------ access
    def get_variables_data(self, ids="all"):
        """{id: dict(tau, v)} (reference message_passing.py:271-276: a copy of the
        variable's node attributes)."""
        st = self._ensure_state()
        vx, vz = st["vx"].cpu().numpy(), st["vz"].cpu().numpy()
        data = {}
        for vid, tau, v in ((self.x_id, self.tau_x, vx), (self.z_id, self.tau_z, vz)):
            if ids == "all" or vid in ids:
                data[vid] = dict(tau=self._out(tau), v=self._out(v))
        return data

    def get_variable_data(self, id):
        data = self.get_variables_data([id])
        if id not in data:
            raise ValueError(f"id={id} not in variables")
        return data[id]

    def _node(self, role):
        return dict(prior=self.prior, x=self.x_var, lin=self.linear, z=self.z_var, lik=self.lik)[role]

    def get_edges_data(self, keys):
        """reference message_passing.py:278-287."""
        st = self._ensure_state()
        a = st["edge_a"].cpu().numpy()
        tau = {"x": self.tau_x, "z": self.tau_z}
        damp = {"e1": "e1", "e3": "e3", "e5": "e5", "e7": "e7"}
        records = []
        for i, (name, role, direction) in enumerate(EDGES):
            s, tgt = EDGE_ENDS[name]
            factor = self._node(s if s in ("prior", "lin", "lik") else tgt)
            values = dict(a=self._out(a[i]), direction=direction, tau=self._out(tau[role]),
                          n_iter=self.n_iter,
                          damping=(self.damp.get(damp.get(name)) or None) if hasattr(self, "damp") else None,
                          A=getattr(self, "A_edges", {}).get(name))
            record = dict(x_id=self._node(role).id, f_id=factor.id)
            for key in keys:
                record[key] = values.get(key)
            records.append(record)
        return records

    def get_nodes_data(self, keys):
        """reference message_passing.py:289-299."""
        data = self.get_variables_data()
        A = getattr(self, "A_nodes", {})
        records = []
        for role in ("prior", "x", "lin", "z", "lik"):
            node = self._node(role)
            is_var = role in ("x", "z")
            values = dict(data[node.id]) if is_var else {}
            values["A"] = A.get(node.id)
            record = dict(id=node.id, type="variable" if is_var else "factor")
            for key in keys:
                record[key] = values.get(key)
            record["n_iter"] = self.n_iter
            records.append(record)
        return records

    # --------------------------------------------------------------- objective
    def update_objective(self):
        """reference message_passing.py:306-328 with node_objective = free energy
        (state_evolution.py:22-23).  Cold path: the averaged log-partitions of the
        prior and the likelihood are device quadratures, the rest scalar closed
        forms."""
        st = self._ensure_state()
        a = st["edge_a"].cpu().numpy()
        A_nodes = np.zeros((5, self.G))
        A_edges = np.zeros((4, self.G))
        with np.errstate(all="ignore"):
            for g, (prior, x_var, lin, z_var, lik) in enumerate(self.chains):
                e = {name: a[i, g] for name, i in EDGE_INDEX.items()}
                tx, tz = self.tau_x[g], self.tau_z[g]
                A_nodes[0, g] = prior.compute_free_energy(e["e8"])
                A_nodes[1, g] = x_var.compute_free_energy(e["e1"] + e["e7"], tx)
                A_nodes[2, g] = lin.compute_free_energy(e["e2"], e["e6"], tx)
                A_nodes[3, g] = z_var.compute_free_energy(e["e3"] + e["e5"], tz)
                A_nodes[4, g] = lik.compute_free_energy(e["e4"], tz)
                for k, (f, b, var, tau) in enumerate((("e1", "e8", x_var, tx), ("e2", "e7", x_var, tx),
                                                      ("e3", "e6", z_var, tz), ("e4", "e5", z_var, tz))):
                    A_edges[k, g] = var.compute_free_energy(e[f] + e[b], tau)
        ids = [self.prior.id, self.x_id, self.linear.id, self.z_id, self.lik.id]
        self.A_nodes = {i: self._out(A_nodes[k]) for k, i in enumerate(ids)}
        self.A_edges = {}
        for k, (f, b) in enumerate((("e1", "e8"), ("e2", "e7"), ("e3", "e6"), ("e4", "e5"))):
            self.A_edges[f] = self.A_edges[b] = self._out(A_edges[k])
        self.A_model = self._out(A_nodes.sum(0) - A_edges.sum(0))

    def entropy(self, update=True):
        """reference state_evolution.py:25-28."""
        if update:
            self.update_objective()
        return -self.A_model
