"""Factor-by-factor EP schedule kept on the device.

reference: tramp/algos/message_passing.py -- `damping="adaptive"` (:151-185) and `update_dA=True`
(:129-149, :241-247) need the local EP objective around every single message, so they cannot run
inside the lock-step sweep of `trb_sweep_run`.  Here the schedule of :249-269 is walked node by node
with all eight messages (a [B], b [B, ld]) resident on the device and every step a kernel launch:

  candidates     prior / likelihood messages by `trb_factor_message`, channel messages by the GEMV
                 kernels + `trb_lin_rescale` + `trb_message_from_posterior`, pass-through copies
  objective      A(target node) - A(variable of the edge): `trb_factor_log_partition` (prior,
                 likelihood), `trb_lin_project` + `trb_lin_log_partition` (channel),
                 `trb_variable_log_partition` (variables)
  step halving   `trb_message_trial` (old + beta (new - old)), the accept mask `dA >= 0` stays a
                 device tensor, `trb_rows_select` keeps the accepted trial of every instance

Nothing is read back while a sweep is enqueued, except one "has every instance accepted?" flag per
trial step of the adaptive damping (the usual case accepts beta = 1: one flag per message) and one
NaN flag per iteration (the reference's check_message); messages and posteriors never leave the
device unless a callback asks for them.  Every instance of a batch takes its own
decisions.  `tramp_b200/algos/factor_schedule.py` is the same schedule evaluated through the numpy
factor API (kept as the cross-check: `ExpectationPropagation.schedule_backend = "host"`).

Chain and edge names:  prior -e1-> x -e2-> lin -e3-> z -e4-> lik,
lik -e5-> z -e6-> lin -e7-> x -e8-> prior.
"""
import logging
import numpy as np

from .. import ops
from .factor_schedule import EDGE_ENDS, OPPOSITE, N_HALVINGS

logger = logging.getLogger(__name__)

ROLE = {"e1": "x", "e2": "x", "e7": "x", "e8": "x", "e3": "z", "e4": "z", "e5": "z", "e6": "z"}
DIRECTION = {"e1": "fwd", "e2": "fwd", "e3": "fwd", "e4": "fwd", "e5": "bwd", "e6": "bwd", "e7": "bwd", "e8": "bwd"}
# what the target node of an edge is made of: a variable (the two messages meeting on it), the
# channel (e2, e6), or a separable factor (its single incoming message)
TARGET = {"e1": ("var", "e1", "e7"), "e7": ("var", "e1", "e7"), "e3": ("var", "e3", "e5"), "e5": ("var", "e3", "e5"),
          "e2": ("lin",), "e6": ("lin",), "e4": ("lik",), "e8": ("prior",)}


class DeviceSchedule:
    """All eight messages on the device + the node-by-node sweep (see the module docstring)."""

    def __init__(self, mp, edges):
        """edges: name -> dict(a=tensor [B], b=tensor [B, ld], n_iter, damping)."""
        self.mp = mp
        self.t = ops.torch()
        self.B, self.N, self.M = mp.B, mp.N, mp.M
        self.lin = mp.linear
        self.msg = {k: dict(a=e["a"].clone(), b=e["b"].clone()) for k, e in edges.items()}
        self.meta = {k: dict(direction=DIRECTION[k], n_iter=int(e.get("n_iter", 0)), damping=e.get("damping"),
                             dA=None, beta=None, A=None) for k, e in edges.items()}
        self.post = {"x": None, "z": None}          # role -> (r [B, ld], v [B])
        self.node_A = {}
        self.y = mp._state["y"]
        self._host_edges = None
        self.old = self.copy_state()

    # ------------------------------------------------------------------ state
    def n_of(self, name):
        return self.N if ROLE[name] == "x" else self.M

    def copy_state(self):
        return ({k: dict(a=m["a"].clone(), b=m["b"].clone()) for k, m in self.msg.items()},
                {k: dict(v) for k, v in self.meta.items()},
                {k: None if p is None else (p[0].clone(), p[1].clone()) for k, p in self.post.items()})

    def restore_state(self, state):
        for k, m in state[0].items():
            self.msg[k]["a"].copy_(m["a"])
            self.msg[k]["b"].copy_(m["b"])
        self.meta = {k: dict(v) for k, v in state[1].items()}
        self.post = {k: None if p is None else (p[0].clone(), p[1].clone()) for k, p in state[2].items()}
        self._host_edges = None

    def set_damping(self, name, value):
        self.meta[name]["damping"] = value
        self._host_edges = None

    def _to_host(self, tensor, n=None):
        x = tensor.cpu().numpy()
        if n is not None:
            x = x[:, :n]
        if self.mp.batched:
            return x
        return float(x[0]) if x.ndim == 1 else x[0]

    @property
    def edges(self):
        """Host view of the messages and their records, in the format of FactorSchedule.edges
        (read-only: built on demand, one device read per message)."""
        if self._host_edges is None:
            out = {}
            for k, m in self.msg.items():
                d = dict(a=self._to_host(m["a"]), b=self._to_host(m["b"], self.n_of(k)))
                for key, val in self.meta[k].items():
                    d[key] = self._to_host(val) if ops.is_tensor(val) else val
                out[k] = d
            self._host_edges = out
        return self._host_edges

    @property
    def variables(self):
        out = {}
        for role, n in (("x", self.N), ("z", self.M)):
            p = self.post[role]
            out[role] = {} if p is None else dict(r=self._to_host(p[0], n), v=self._to_host(p[1]))
        return out

    def mirror_into(self, st):
        """Copy the messages and posteriors into the state of the lock-step sweep (device to device),
        so that get_variables_data / snapshots / a later device-path warm start see them."""
        for i, name in enumerate(("e1", "e2", "e3", "e4", "e5", "e6", "e7", "e8")):
            st["edge_a"][i].copy_(self.msg[name]["a"])
        for buf, name in (("b1", "e1"), ("b3", "e3"), ("b5", "e5"), ("b7", "e7")):
            st[buf].copy_(self.msg[name]["b"])
        for role, rk, vk in (("x", "rx", "vx"), ("z", "rz", "vz")):
            if self.post[role] is not None:
                st[rk].copy_(self.post[role][0])
                st[vk].copy_(self.post[role][1])

    def aliases_hold(self):
        """e2 == e1, e4 == e3, e6 == e5, e8 == e7 (needed to hand over to the lock-step sweep)."""
        t = self.t
        return all(bool(t.equal(self.msg[c]["a"], self.msg[s]["a"]) and t.equal(self.msg[c]["b"], self.msg[s]["b"]))
                   for c, s in (("e2", "e1"), ("e4", "e3"), ("e6", "e5"), ("e8", "e7")))

    # -------------------------------------------------------------- objective
    def _project(self, name, b):
        """tz = V_R^T b (x side: e2) or tx = U_R^T b (z side: e6)."""
        lin = self.lin
        if ROLE[name] == "x":
            return ops.lin_project(lin.Vt, lin.R, lin.Nz, b, self.B)
        return ops.lin_project(lin.Ut, lin.R, lin.Nx, b, self.B)

    def _lin_objective(self, e2, e6, tz=None, tx=None):
        lin = self.lin
        tz = self._project("e2", e2["b"]) if tz is None else tz
        tx = self._project("e6", e6["b"]) if tx is None else tx
        bz2 = ops.row_dot(e2["b"], e2["b"], self.N) if lin.R < lin.Nz else None
        return ops.lin_log_partition(lin.s, lin.s2, lin.Nz, e2["a"], e6["a"], tz, tx, bz2)

    def _factor_objective(self, which, m):
        mp = self.mp
        if which == "prior":
            return ops.factor_log_partition(mp.prior._trb_factor(), m["a"], m["b"], None, self.N, False, False)
        return ops.factor_log_partition(mp.lik._trb_factor(), m["a"], m["b"], self.y, self.M, False, False)

    def _var_objective(self, m1, m2, n):
        return ops.variable_log_partition(m1["a"], m1["b"], m2["a"], m2["b"], n)

    def objective_around(self, name, data=None, cache=None):
        """A(target node) - A(variable of the edge) [B] with edge `name` carrying `data`
        (:137-149 / :160-174).  cache: the projection of the channel's OTHER message, which the
        trials of one line search share."""
        cur = lambda k: data if (data is not None and k == name) else self.msg[k]   # noqa: E731
        kind = TARGET[name]
        if kind[0] == "var":
            A_target = self._var_objective(cur(kind[1]), cur(kind[2]), self.n_of(name))
        elif kind[0] == "lin":
            other = "e6" if name == "e2" else "e2"
            if cache is not None and other not in cache:
                cache[other] = self._project(other, self.msg[other]["b"])
            fixed = None if cache is None else cache[other]
            A_target = self._lin_objective(cur("e2"), cur("e6"), tz=fixed if other == "e2" else None,
                                           tx=fixed if other == "e6" else None)
        else:
            A_target = self._factor_objective(kind[0], cur(name))
        A_edge = self._var_objective(cur(name), self.msg[OPPOSITE[name]], self.n_of(name))
        return A_target - A_edge

    # ------------------------------------------------------------- candidates
    def _channel_posterior(self, direction):
        """Posterior mean and variance of the channel's output x-side (0) or input z-side (1)."""
        lin, B = self.lin, self.B
        e2, e6 = self.msg["e2"], self.msg["e6"]
        tz = ops.lin_project(lin.Vt, lin.R, lin.Nz, e2["b"], B)
        tx = ops.lin_project(lin.Ut, lin.R, lin.Nx, e6["b"], B)
        coef, v = ops.lin_rescale(direction, B, lin.R, lin.Nz, lin.Nx, lin.rank, lin.s, lin.s2,
                                  e2["a"], e6["a"], tz, tx)
        if direction == 0:
            return ops.lin_expand(lin.Ut, lin.R, lin.Nx, coef, B), v
        null = lin.R < lin.Nz
        r = ops.lin_expand(lin.Vt, lin.R, lin.Nz, coef, B, add=e2["b"] if null else None,
                           add_div=e2["a"] if null else None)
        return r, v

    def candidate(self, name):
        """The undamped new message of edge `name` from the current messages (:249-269)."""
        mp, t = self.mp, self.t
        if name in ("e2", "e4", "e6", "e8"):                      # SISOVariable pass-through
            src = self.msg[{"e2": "e1", "e4": "e3", "e6": "e5", "e8": "e7"}[name]]
            return dict(a=src["a"].clone(), b=src["b"].clone())
        if name in ("e1", "e5"):                                  # separable factor
            factor, src, y, n = ((mp.prior._trb_factor(), self.msg["e8"], None, self.N) if name == "e1"
                                 else (mp.lik._trb_factor(), self.msg["e4"], self.y, self.M))
            a_new = t.zeros_like(src["a"])
            b_new = t.zeros_like(src["b"])
            flags = t.zeros(self.B, dtype=t.int32, device=src["b"].device)
            ops.factor_message(factor, src["a"], src["b"], y, n, a_new, b_new, damping=0.0, flags=flags)
            return dict(a=a_new, b=b_new)
        lin = self.lin
        if name == "e3":                                          # channel -> z (forward message)
            r, v = self._channel_posterior(0)
            a_new, b_new = ops.message_from_posterior(r, v, self.msg["e6"]["a"], self.msg["e6"]["b"], self.M,
                                                      lin.AMIN, lin.AMAX)
        else:                                                     # e7: channel -> x (backward message)
            r, v = self._channel_posterior(1)
            a_new, b_new = ops.message_from_posterior(r, v, self.msg["e2"]["a"], self.msg["e2"]["b"], self.N,
                                                      lin.AMIN, lin.AMAX)
        return dict(a=a_new, b=b_new)

    # ---------------------------------------------------------------- damping
    def adaptive_damping(self, name, new):
        """:151-185, every instance on its own: the first beta = 2^-n whose local objective does not
        decrease is kept; none after ten trials: the old message, dA = 0, beta = 0."""
        t, n = self.t, self.n_of(name)
        old = self.msg[name]
        cache = {}
        A_old = self.objective_around(name, cache=cache)
        kept = dict(a=old["a"].clone(), b=old["b"].clone())
        dA_kept, beta_kept = t.zeros_like(A_old), t.zeros_like(A_old)
        accepted = t.zeros(self.B, dtype=t.bool, device=A_old.device)
        trial = dict(a=t.empty_like(old["a"]), b=t.zeros_like(old["b"]))
        for k in range(N_HALVINGS):
            beta = 1 / 2**k
            ops.message_trial(old["a"], old["b"], new["a"], new["b"], n, beta, trial["a"], trial["b"])
            dA = self.objective_around(name, trial, cache=cache) - A_old
            take = (dA >= 0) & ~accepted                         # NaN never passes, as in the reference
            ops.rows_select(take.to(t.int32), trial["a"], trial["b"], kept["a"], kept["b"], n)
            dA_kept = t.where(take, dA, dA_kept)
            beta_kept = t.where(take, t.full_like(dA, beta), beta_kept)
            accepted = accepted | take
            if bool(accepted.all().item()):                       # one flag per trial; the usual case stops at beta = 1
                break
        return kept, dA_kept, beta_kept

    def constant_damping(self, name, new):
        """:119-127."""
        d = self.meta[name].get("damping")
        if not d:
            return new
        old = self.msg[name]
        out = dict(a=self.t.empty_like(old["a"]), b=self.t.zeros_like(old["b"]))
        ops.message_trial(new["a"], new["b"], old["a"], old["b"], self.n_of(name), float(d), out["a"], out["b"])
        return out                                               # new + d (old - new) = d old + (1 - d) new

    # ------------------------------------------------------------------ sweep
    def emit(self, name):
        mp, t = self.mp, self.t
        data = self.candidate(name)
        self._nan = self._nan | t.isnan(data["a"]).any() | t.isnan(data["b"]).any()
        meta = self.meta[name]
        if mp.damping:
            if mp.adaptive_damping:
                if mp.n_iter > 0:
                    data, meta["dA"], meta["beta"] = self.adaptive_damping(name, data)
            else:
                data = self.constant_damping(name, data)
        meta["n_iter"] += 1
        if mp.update_dA:
            if mp.n_iter > 0:
                meta["dA"] = self.objective_around(name, data) - self.objective_around(name)
            else:
                meta["dA"] = t.zeros(self.B, dtype=t.float64, device=data["a"].device)
        self.msg[name]["a"].copy_(data["a"])
        self.msg[name]["b"].copy_(data["b"])

    def sweep(self):
        """One iteration: forward pass, backward pass, update_variables (:249-269)."""
        t = self.t
        self._nan = t.zeros((), dtype=t.bool, device=self.msg["e1"]["a"].device)
        self._host_edges = None
        for name in ("e1", "e2", "e3", "e4", "e5", "e6", "e7", "e8"):
            self.emit(name)
        if bool(self._nan.item()):                               # check_message (:187-209), once per iteration
            logger.warning("restoring old message dag")
            self.restore_state(self.old)
            raise ValueError("EP message a or b is nan")
        for role, f, bwd, n in (("x", "e1", "e7", self.N), ("z", "e3", "e5", self.M)):
            self.post[role] = ops.posterior_rv(self.msg[f]["a"], self.msg[f]["b"], self.msg[bwd]["a"],
                                               self.msg[bwd]["b"], n)

    def update_objective(self):
        """:306-328 on the un-aliased device messages; returns A_model (float, or [B] for a batch)."""
        m = self.msg
        A = dict(prior=self._factor_objective("prior", m["e8"]), x=self._var_objective(m["e1"], m["e7"], self.N),
                 lin=self._lin_objective(m["e2"], m["e6"]), z=self._var_objective(m["e3"], m["e5"], self.M),
                 lik=self._factor_objective("lik", m["e4"]))
        self.node_A = {k: self._to_host(v) for k, v in A.items()}
        total = sum(A.values())
        for name in ("e1", "e2", "e3", "e4"):
            Ae = self._var_objective(m[name], m[OPPOSITE[name]], self.n_of(name))
            self.meta[name]["A"] = self.meta[OPPOSITE[name]]["A"] = self._to_host(Ae)
            total = total - Ae
        self._host_edges = None
        return self._to_host(total)
