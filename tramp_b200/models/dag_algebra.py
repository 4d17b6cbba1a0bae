"""DAG algebra (reference tramp/models/dag_algebra.py), re-stated on a small
ordered-adjacency graph instead of networkx 1.x.  Supports what the EP hot path
needs: `@` composition with placeholder splicing (:115-132), `to_model()`
(:140-142) and `to_observed()` (:243-291)."""
from ..base import Variable, Factor, ReprMixin
from ..channels import GaussianChannel, AbsChannel, SgnChannel
from ..likelihoods import GaussianLikelihood, AbsLikelihood, SgnLikelihood


def channel2likelihood(channel, y, y_name):
    """reference dag_algebra.py:21-40 (in-scope channels)."""
    if isinstance(channel, GaussianChannel):
        return GaussianLikelihood(y=y, y_name=y_name, var=channel.var)
    if isinstance(channel, AbsChannel):
        return AbsLikelihood(y=y, y_name=y_name)
    if isinstance(channel, SgnChannel):
        return SgnLikelihood(y=y, y_name=y_name)
    raise NotImplementedError(f"cannot convert {channel} to likelihood")


class PlaceHolder(ReprMixin):
    def __init__(self):
        self.repr_init()

    def math(self):
        return r"$\emptyset$"


class RootPlaceHolder(PlaceHolder):
    n_prev = 0
    n_next = 1


class LeafPlaceHolder(PlaceHolder):
    n_prev = 1
    n_next = 0


class Graph:
    """Minimal directed graph with insertion-ordered nodes and edges."""

    def __init__(self):
        self.succ = {}
        self.pred = {}
        self.node = {}

    def add_node(self, n, **attr):
        if n not in self.succ:
            self.succ[n], self.pred[n], self.node[n] = [], [], {}
        self.node[n].update(attr)

    def add_edge(self, u, v):
        self.add_node(u)
        self.add_node(v)
        if v not in self.succ[u]:
            self.succ[u].append(v)
            self.pred[v].append(u)

    def add_edges_from(self, edges):
        for u, v in edges:
            self.add_edge(u, v)

    def remove_node(self, n):
        for v in self.succ.pop(n):
            self.pred[v].remove(n)
        for u in self.pred.pop(n):
            self.succ[u].remove(n)
        self.node.pop(n)

    def nodes(self):
        return list(self.succ)

    def edges(self):
        return [(u, v) for u in self.succ for v in self.succ[u]]

    def predecessors(self, n):
        return list(self.pred[n])

    def successors(self, n):
        return list(self.succ[n])

    def copy(self):
        g = Graph()
        for n in self.succ:
            g.add_node(n, **self.node[n])
        g.add_edges_from(self.edges())
        return g

    def topological_sort(self):
        indeg = {n: len(self.pred[n]) for n in self.succ}
        ready = [n for n in self.succ if indeg[n] == 0]
        order = []
        while ready:
            n = ready.pop(0)
            order.append(n)
            for v in self.succ[n]:
                indeg[v] -= 1
                if indeg[v] == 0:
                    ready.append(v)
        if len(order) != len(self.succ):
            raise ValueError("graph has a cycle")
        return order


def check_dag(dag):
    """reference dag_algebra.py:59-77."""
    if not isinstance(dag, Graph):
        raise ValueError(f"dag {dag} not a DAG")
    dag.topological_sort()
    for node in dag.nodes():
        n_prev, n_next = len(dag.predecessors(node)), len(dag.successors(node))
        if n_prev != node.n_prev:
            raise ValueError(f"node {node} has {n_prev} predecessors but should have {node.n_prev}")
        if n_next != node.n_next:
            raise ValueError(f"node {node} has {n_next} successors but should have {node.n_next}")


def to_dag(node):
    """reference dag_algebra.py:80-87."""
    dag = Graph()
    dag.add_node(node)
    for _ in range(node.n_next):
        dag.add_edge(node, LeafPlaceHolder())
    for _ in range(node.n_prev):
        dag.add_edge(RootPlaceHolder(), node)
    return dag


class DAG():
    """reference dag_algebra.py:90-142."""

    def __init__(self, dag):
        if not isinstance(dag, Graph):
            dag = to_dag(dag)
        check_dag(dag)
        self.dag = dag
        nodes = dag.topological_sort()
        self._leafs_ph = [n for n in nodes if isinstance(n, LeafPlaceHolder)]
        self._roots_ph = [n for n in nodes if isinstance(n, RootPlaceHolder)]

    def __add__(self, other):
        raise NotImplementedError(
            "`+` (tree-structured models) is outside the EP hot path of tramp_b200; "
            "only chains built with `@` are supported")

    def __matmul__(self, other):
        if not isinstance(other, DAG):
            other = DAG(other)
        dag = Graph()
        dag.add_edges_from(self.dag.edges())
        dag.add_edges_from(other.dag.edges())
        for n in self.dag.nodes() + other.dag.nodes():
            dag.add_node(n)
        # dag surgery: splice each leaf placeholder of self with a root placeholder of other
        for leaf, root in zip(self._leafs_ph, other._roots_ph):
            leaf_predecessors = self.dag.predecessors(leaf)
            root_successors = other.dag.successors(root)
            assert len(leaf_predecessors) == 1
            assert len(root_successors) == 1
            dag.remove_node(leaf)
            dag.remove_node(root)
            dag.add_edge(leaf_predecessors[0], root_successors[0])
        return DAG(dag)

    def to_model_dag(self):
        return ModelDAG(self.dag)

    def to_model(self):
        from .base_model import Model
        return Model(self.to_model_dag())


def check_model_dag(dag):
    """reference dag_algebra.py:216-233."""
    if not isinstance(dag, Graph):
        raise ValueError(f"dag {dag} not a DAG")
    for node in dag.nodes():
        if not (isinstance(node, Factor) or isinstance(node, Variable)):
            raise ValueError(f"node {node} should be a Factor or Variable")
        opposite_class = Factor if isinstance(node, Variable) else Variable
        for other in dag.predecessors(node) + dag.successors(node):
            if not isinstance(other, opposite_class):
                raise ValueError(f"neighbour {other} of {node} must be a {opposite_class}")


class ModelDAG(DAG):
    """reference dag_algebra.py:236-291."""

    def __init__(self, dag):
        if isinstance(dag, Variable) or isinstance(dag, Factor):
            dag = to_dag(dag)
        check_model_dag(dag)
        super().__init__(dag)

    def to_observed(self, observations):
        """ModelDAG with observed variables: observations = {id: observation}."""
        observed_ids = observations.keys()

        def is_observed(node):
            return isinstance(node, Variable) and node.id in observed_ids

        def is_likelihood(node):
            if not isinstance(node, Factor):
                return False
            return any(v.id in observed_ids for v in self.dag.successors(node))

        dag = Graph()
        for source, target in self.dag.edges():
            if is_observed(target):
                if target.n_next != 0:
                    raise ValueError(f"{target} not a leaf")
            elif is_likelihood(target):
                ids = [v.id for v in self.dag.successors(target) if v.id in observed_ids]
                if len(ids) != 1:
                    raise ValueError(f"cannot convert {target} to likelihood")
                likelihood = channel2likelihood(target, y=observations[ids[0]], y_name=ids[0])
                dag.add_edge(source, likelihood)
            else:
                dag.add_edge(source, target)
        return ModelDAG(dag)
