"""Variable arities on the EP hot path (reference tramp/variables/sub_variables.py).

On the device a SISOVariable is an exact pass-through: its two outgoing edges
alias the incoming ones (reference sub_variables.py:16-31), so it costs no
kernel."""
from ..base import Variable


class SISOVariable(Variable):
    """Single-in single-out variable (reference sub_variables.py:16-31)."""

    def __init__(self, id):
        super().__init__(id=id, n_prev=1, n_next=1)

    def forward_message(self, message):
        "pass message from previous factor k to next factor l, unchanged"
        from ..base import filter_message
        (k, _, dk), = filter_message(message, "fwd")
        (l, _, dl), = filter_message(message, "bwd")
        return [(self, l, dict(a=dk["a"], b=dk["b"], direction="fwd"))]

    def backward_message(self, message):
        "pass message from next factor l to previous factor k, unchanged"
        from ..base import filter_message
        (k, _, dk), = filter_message(message, "fwd")
        (l, _, dl), = filter_message(message, "bwd")
        return [(self, k, dict(a=dl["a"], b=dl["b"], direction="bwd"))]


    def forward_state_evolution(self, message):
        "State Evolution: the precision passes through unchanged (reference sub_variables.py:33-37)"
        from ..base import filter_message
        (k, _, dk), = filter_message(message, "fwd")
        (l, _, dl), = filter_message(message, "bwd")
        return [(self, l, dict(a=dk["a"], direction="fwd"))]

    def backward_state_evolution(self, message):
        "reference sub_variables.py:39-43"
        from ..base import filter_message
        (k, _, dk), = filter_message(message, "fwd")
        (l, _, dl), = filter_message(message, "bwd")
        return [(self, k, dict(a=dl["a"], direction="bwd"))]


class SILeafVariable(Variable):
    """Observed leaf (reference sub_variables.py:52-55)."""

    def __init__(self, id):
        super().__init__(id=id, n_prev=1, n_next=0)


class SORootVariable(Variable):
    def __init__(self, id):
        super().__init__(id=id, n_prev=0, n_next=1)
