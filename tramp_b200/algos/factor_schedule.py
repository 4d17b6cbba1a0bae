"""Factor-by-factor EP schedule driven from the host.

reference: tramp/algos/message_passing.py -- the options that need the EP
objective after EVERY single message, `damping="adaptive"` (:151-185) and
`update_dA=True` (:129-149, :241-247), cannot run inside the lock-step device
sweep of `trb_sweep_run`.  They take this path instead: the schedule of
:249-269 is walked node by node on the host, and every factor evaluation
(`compute_*_message`, `compute_log_partition`) still runs in the CUDA kernels
through the factor API of tramp_b200.priors / likelihoods / channels.  Only
the per-message decisions (`dA >= 0`, NaN checks) are host arithmetic on a few
scalars per instance.

Batched models run all their instances through every factor evaluation at once
(`a` is then an array [B], `b` an array [B, n]); the decisions stay PER INSTANCE:
each instance halves its own step until its own objective stops decreasing, so a
batched run equals the per-instance runs.

Chain and edge names:  prior -e1-> x -e2-> lin -e3-> z -e4-> lik,
lik -e5-> z -e6-> lin -e7-> x -e8-> prior.
"""
import logging
import numpy as np

logger = logging.getLogger(__name__)

EDGE_ENDS = {"e1": ("prior", "x"), "e2": ("x", "lin"), "e3": ("lin", "z"), "e4": ("z", "lik"),
             "e5": ("lik", "z"), "e6": ("z", "lin"), "e7": ("lin", "x"), "e8": ("x", "prior")}
OPPOSITE = {"e1": "e8", "e8": "e1", "e2": "e7", "e7": "e2", "e3": "e6", "e6": "e3", "e4": "e5", "e5": "e4"}
IN_EDGES = {"prior": ("e8",), "x": ("e1", "e7"), "lin": ("e2", "e6"), "z": ("e3", "e5"), "lik": ("e4",)}
FORWARD_ORDER = ("prior", "x", "lin", "z", "lik")
VARIABLES = ("x", "z")
N_HALVINGS = 10   # message_passing.py:168


def _per_instance(mask, like):
    """Boolean [B] (or scalar) mask shaped to select whole rows of `like`."""
    mask = np.asarray(mask)
    return mask[..., None] if np.ndim(like) > mask.ndim else mask


class FactorSchedule:
    """Host-side message state (all eight edges, not aliased) + the node-by-node sweep."""

    def __init__(self, mp, edges):
        self.mp = mp
        self.nodes = dict(prior=mp.prior, x=mp.x_var, lin=mp.linear, z=mp.z_var, lik=mp.lik)
        self.edges = edges            # name -> dict(a, b, direction, n_iter, damping, ...)
        self.variables = {"x": {}, "z": {}}
        self.node_A = {}
        self.old = self.copy_state()

    # ------------------------------------------------------------------ state
    def copy_state(self):
        return ({k: dict(d) for k, d in self.edges.items()},
                {k: dict(d) for k, d in self.variables.items()})

    def restore_state(self, state):
        self.edges = {k: dict(d) for k, d in state[0].items()}
        self.variables = {k: dict(d) for k, d in state[1].items()}

    def message(self, node, replace=None, data=None):
        """Incoming messages of `node` as the reference's [(source, target, data)];
        the data of edge `replace` is substituted (create_message, :47-52)."""
        out = []
        for name in IN_EDGES[node]:
            s, t = EDGE_ENDS[name]
            out.append((self.nodes[s], self.nodes[t], data if name == replace else self.edges[name]))
        return out

    def edge_pair(self, name, data=None):
        """The two opposite messages living on one model edge (:139-142)."""
        s, t = EDGE_ENDS[name]
        opp = OPPOSITE[name]
        return [(self.nodes[s], self.nodes[t], self.edges[name] if data is None else data),
                (self.nodes[t], self.nodes[s], self.edges[opp])]

    # -------------------------------------------------------------- objective
    def objective_around(self, name, data=None):
        """A(target node) - A(variable of the edge) with edge `name` carrying `data`
        (:137-149 / :160-174)."""
        s, t = EDGE_ENDS[name]
        variable = t if t in VARIABLES else s
        m_target = self.message(t, replace=name if data is not None else None, data=data)
        with np.errstate(all="ignore"):
            A_target = self.mp.node_objective(self.nodes[t], m_target)
            A_edge = self.mp.node_objective(self.nodes[variable], self.edge_pair(name, data))
        return A_target - A_edge

    def compute_dA(self, name, data):
        if self.mp.n_iter == 0:
            return 0
        return self.objective_around(name, data) - self.objective_around(name)

    def adaptive_damping(self, name, data):
        """:151-185: halve the step until the local objective does not decrease -- every
        instance of a batch on its own (all of them are evaluated at every trial step; an
        instance keeps the first step size that its own objective accepts)."""
        if self.mp.n_iter == 0:
            return data
        old = self.edges[name]
        step = {k: data[k] - old[k] for k in ("a", "b")}
        A_old = self.objective_around(name)
        accepted = np.zeros(np.shape(A_old), dtype=bool)
        # not accepted after N_HALVINGS trials: the old message, dA = 0, beta = 0 (:182-185)
        kept = {k: np.array(old[k], dtype=float) for k in ("a", "b")}
        dA_kept, beta_kept = np.zeros(np.shape(A_old)), np.zeros(np.shape(A_old))
        for n in range(N_HALVINGS):
            beta = 1 / 2**n
            trial = dict(data)
            for k in ("a", "b"):
                trial[k] = old[k] + beta * step[k]
            dA = self.objective_around(name, trial) - A_old
            take = np.logical_and(np.asarray(dA) >= 0, ~accepted)       # NaN never passes, as in the reference
            for k in ("a", "b"):
                kept[k] = np.where(_per_instance(take, kept[k]), trial[k], kept[k])
            dA_kept = np.where(take, dA, dA_kept)
            beta_kept = np.where(take, beta, beta_kept)
            accepted = accepted | take
            if accepted.all():
                break
        new = dict(data)
        if self.mp.batched:
            new.update(a=kept["a"], b=kept["b"], dA=dA_kept, beta=beta_kept)
        else:
            new.update(a=float(kept["a"]), b=kept["b"], dA=float(dA_kept), beta=float(beta_kept))
        return new

    def constant_damping(self, name, data):
        """:119-127."""
        d = self.edges[name].get("damping")
        if not d:
            return data
        old = self.edges[name]
        new = dict(data)
        for k in ("a", "b"):
            new[k] = d * old[k] + (1 - d) * data[k]
        return new

    # ------------------------------------------------------------------ sweep
    def check(self, name, data):
        """:187-209."""
        s, t = EDGE_ENDS[name]
        sid, tid = self.nodes[s].id, self.nodes[t].id
        if np.any(np.isnan(data["a"])):
            logger.warning("restoring old message dag")
            self.restore_state(self.old)
            raise ValueError(f"{sid}->{tid} a is nan")
        if np.any(np.asarray(data["a"]) < 0):
            logger.warning(f"{sid}->{tid} negative a {data['a']}")
        if np.isnan(data["b"]).any():
            logger.warning("restoring old message dag")
            self.restore_state(self.old)
            raise ValueError(f"{sid}->{tid} b is nan")

    def emit(self, node, new_message):
        mp = self.mp
        for source, target, data in new_message:
            name = next(k for k, (s, t) in EDGE_ENDS.items()
                        if self.nodes[s] is source and self.nodes[t] is target)
            data = dict(data)
            data["a"] = (np.asarray(data["a"], dtype=np.float64).reshape(mp.B) if mp.batched
                         else float(np.asarray(data["a"])))
            data["b"] = np.asarray(data["b"], dtype=np.float64)
            self.check(name, data)
            if mp.damping:
                data = (self.adaptive_damping(name, data) if mp.adaptive_damping
                        else self.constant_damping(name, data))
            data["n_iter"] = self.edges[name]["n_iter"] + 1
            if mp.update_dA:
                data["dA"] = self.compute_dA(name, data)
            self.edges[name].update(data)

    def sweep(self):
        """One iteration: forward pass, backward pass, update_variables (:249-269)."""
        mp = self.mp
        with np.errstate(all="ignore"):
            for node in FORWARD_ORDER:
                self.emit(node, mp.forward(self.nodes[node], self.message(node)))
            for node in reversed(FORWARD_ORDER):
                self.emit(node, mp.backward(self.nodes[node], self.message(node)))
            for v in VARIABLES:
                self.variables[v] = mp.update(self.nodes[v], self.message(v))

    def update_objective(self):
        """:306-328 on the un-aliased host messages."""
        with np.errstate(all="ignore"):
            for node in FORWARD_ORDER:
                self.node_A[node] = self.mp.node_objective(self.nodes[node], self.message(node))
            for name in ("e1", "e2", "e3", "e4"):
                s, t = EDGE_ENDS[name]
                variable = t if t in VARIABLES else s
                A = self.mp.node_objective(self.nodes[variable], self.edge_pair(name))
                self.edges[name]["A"] = A
                self.edges[OPPOSITE[name]]["A"] = A
        return (sum(self.node_A.values()) - sum(self.edges[n]["A"] for n in ("e1", "e2", "e3", "e4")))
