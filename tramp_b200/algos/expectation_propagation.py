"""ExpectationPropagation (reference tramp/algos/expectation_propagation.py:5-32)."""
import numpy as np

from .message_passing import MessagePassing
from .callbacks import EarlyStoppingEP


def _variable_log_partition(ax, bx):
    """reference base.py:146-150, vectorised over a leading batch dimension."""
    ax = np.asarray(ax, dtype=float)
    bx = np.asarray(bx, dtype=float)
    a = ax[..., None] if bx.ndim > ax.ndim else ax
    with np.errstate(all="ignore"):
        logZ = 0.5 * np.sum(bx**2 / a + np.log(2 * np.pi / a), axis=-1)
    return np.where(ax <= 0, np.inf, logZ) if ax.ndim else (np.inf if ax <= 0 else float(logZ))


class ExpectationPropagation(MessagePassing):
    """EP = the message-passing driver with (a, b) messages: a node answers with its
    moment-matched messages, a variable with its posterior (r, v), and the objective is
    the sum of the log-partitions (reference expectation_propagation.py:5-32).  On the
    device-resident path these bindings are what `trb_sweep_run` implements; they are
    called node by node only by the host-driven schedule (adaptive damping, update_dA)."""

    def __init__(self, model):
        model.init_shapes()                 # (a, b) messages need the shapes of the variables
        MessagePassing.__init__(self, model, message_keys=["a", "b"])
        self.default_stopping = EarlyStoppingEP()

    # -- what the driver asks of a node ------------------------------------------------
    @staticmethod
    def forward(node, message):
        return node.forward_message(message)

    @staticmethod
    def backward(node, message):
        return node.backward_message(message)

    @staticmethod
    def update(variable, message):
        return dict(zip(("r", "v"), variable.posterior_rv(message)))

    @staticmethod
    def node_objective(node, message):
        return node.log_partition(message)

    # -- objective -----------------------------------------------------------------------
    def log_evidence(self, update=True):
        "A_model = ln Z of the EP approximation (recomputed from the current messages by default)"
        if update:
            self.update_objective()
        return self.A_model

    def surprisal(self, update=True):
        return -self.log_evidence(update)

    def update_objective(self):
        """reference message_passing.py:306-328: A_model = sum_nodes A - sum_fwd-edges A.
        Cold path: the factor log-partitions run on the device, the (tiny)
        variable terms on the host."""
        if self._host is not None:      # un-aliased host messages (adaptive damping / update_dA)
            self.A_model = self._host.update_objective()
            ids = dict(prior=self.prior.id, x=self.x_id, lin=self.linear.id, z=self.z_id, lik=self.lik.id)
            self.A_nodes = {ids[k]: v for k, v in self._host.node_A.items()}
            self.A_edges = {n: self._host.edges[n]["A"] for n in ("e1", "e2", "e3", "e4")}
            return
        E = {name: self._edge(name) for name in ("e1", "e2", "e3", "e4", "e5", "e6", "e7", "e8")}
        A = {}
        A[self.prior.id] = self.prior.compute_log_partition(*E["e8"])
        A[self.x_id] = _variable_log_partition(E["e1"][0] + E["e7"][0], E["e1"][1] + E["e7"][1])
        A[self.linear.id] = self.linear.compute_log_partition(E["e2"][0], E["e2"][1],
                                                              E["e6"][0], E["e6"][1])
        A[self.z_id] = _variable_log_partition(E["e3"][0] + E["e5"][0], E["e3"][1] + E["e5"][1])
        A[self.lik.id] = self.lik.compute_log_partition(E["e4"][0], E["e4"][1], self.lik.y)
        self.A_nodes = A
        pairs = [("e1", "e8"), ("e2", "e7"), ("e3", "e6"), ("e4", "e5")]
        self.A_edges = {f: _variable_log_partition(E[f][0] + E[b][0], E[f][1] + E[b][1])
                        for f, b in pairs}
        self.A_edge_by_name = {}
        for f, b in pairs:
            self.A_edge_by_name[f] = self.A_edge_by_name[b] = self.A_edges[f]
        self.A_model = sum(A.values()) - sum(self.A_edges.values())
