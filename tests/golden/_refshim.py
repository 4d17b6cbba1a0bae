"""Shims that let the UNMODIFIED reference (sphinxteam/tramp, /root/reference)
import and run in this image.  Used ONLY by tests/golden/make_golden.py, which
runs in the build container; nothing here ships with, or is imported by,
tramp_b200.

* The reference needs networkx<2 (setup.py:13); this image has 3.x.  `install()`
  registers a small module under the name "networkx" exposing the 1.x behaviour
  the EP path relies on (algos/message_passing.py:101-102,212-232,251-269,356;
  models/dag_algebra.py:62-132,199-212,273-291; models/base_model.py:34-35,
  75-124): list-returning accessors, `.node`, dict-copying `copy()`.
* tramp.experiments / tramp.checks import matplotlib.pyplot at module scope
  (experiments/plots.py:3, checks/check_gradients.py:3); an empty stub module is
  registered if matplotlib is missing.
"""
import sys
import types


def _make_networkx_veneer():
    import networkx as real

    class DiGraph(real.DiGraph):
        @property
        def node(self):
            return self._node

        # networkx 3 implements nodes / edges / in_edges as cached properties that
        # store a view object in the instance __dict__, which would shadow these
        # methods from the second call on; drop the cached view after each use.
        def nodes(self, data=False):
            out = list(super().nodes(data=data))
            self.__dict__.pop("nodes", None)
            return out

        def edges(self, nbunch=None, data=False):
            out = list(super().edges(nbunch, data=data))
            self.__dict__.pop("edges", None)
            return out

        def in_edges(self, nbunch=None, data=False):
            out = list(super().in_edges(nbunch, data=data))
            self.__dict__.pop("in_edges", None)
            return out

        def predecessors(self, n):
            return list(super().predecessors(n))

        def successors(self, n):
            return list(super().successors(n))

        def reverse(self, copy=True):
            H = DiGraph()
            H.add_nodes_from((n, dict(d)) for n, d in self._node.items())
            H.add_edges_from(
                (v, u, dict(d)) for u, nb in self._adj.items() for v, d in nb.items())
            return H

        def copy(self):
            # 1.x copy() is a deepcopy; copying each attribute dict is all the
            # EP driver needs (messages are replaced, never mutated in place).
            H = DiGraph()
            H.add_nodes_from((n, dict(d)) for n, d in self._node.items())
            H.add_edges_from(
                (u, v, dict(d)) for u, nb in self._adj.items() for v, d in nb.items())
            return H

    veneer = types.ModuleType("networkx")
    veneer.DiGraph = DiGraph
    veneer.topological_sort = lambda G: list(real.topological_sort(G))
    veneer.is_directed_acyclic_graph = real.is_directed_acyclic_graph
    # 1.x freeze() blocks structural edits only; attribute dicts stay mutable,
    # which message_passing.py:241-247 depends on.
    veneer.freeze = lambda G: G
    veneer.__real__ = real
    return veneer


def install(reference_root="/root/reference"):
    """Make `import tramp` resolve to the unmodified reference."""
    if "tramp" in sys.modules:
        return
    sys.modules["networkx"] = _make_networkx_veneer()
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        mpl.rc = lambda *a, **k: None
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
