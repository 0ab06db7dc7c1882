"""Array-backed symmetric matrix + the ``matrix_de_novo`` drop-in.

``SymMatrix`` keeps the API of /root/reference/src/phamclust/matrix.py:33-405
(the object handed to the unchanged host clustering / heatmap code, SURVEY.md
3.5) but stores the packed upper triangle the device produces: int32
micro-units (``round(w, 6) * 1e6`` exactly, matrix.py:323), so a 50k-genome
matrix is 5 GB of int32 instead of 1.25e9 Python dict entries.

``matrix_de_novo(genomes, func, cpus, as_distance=True)`` has the reference's
signature (matrix.py:432); ``func`` must be one of the six metric functions of
``phamclust_b200.metrics`` -- the whole upper triangle is computed by
libphamclust_b200 in one call instead of one joblib task per pair.
"""
import logging

import numpy as np

from phamclust_b200 import _native
from phamclust_b200.packing import PackedTable, pack_genomes
from phamclust_b200.statistics import average

_UNSET = -1


def _tri_offsets(n):
    """Start of every row in the packed upper triangle (length n+1)."""
    i = np.arange(n + 1, dtype=np.int64)
    return i * n - (i * (i - 1)) // 2


class SymMatrix:
    """Undirected weighted graph over named nodes; drop-in for the reference's
    dict-of-dicts class.  Weights are exact 6-decimal values."""

    def __init__(self, nodes, is_distance=False, _micro=None):
        self._nodes = nodes                       # traversal order (mutable, like upstream)
        self._store = list(nodes)                 # storage order (fixed)
        self._index = {name: k for k, name in enumerate(self._store)}
        n = len(self._store)
        self._offsets = _tri_offsets(n)
        if _micro is None:
            _micro = np.full(n * (n + 1) // 2, _UNSET, dtype=np.int32)
        self._micro = _micro
        self._is_distance = is_distance
        self._locked = False

    # -- construction from device output ---------------------------------------
    @classmethod
    def from_micro(cls, nodes, micro, is_distance):
        n = len(nodes)
        micro = np.ascontiguousarray(micro, dtype=np.int32)
        if micro.size != n * (n + 1) // 2:
            raise ValueError(f"need {n * (n + 1) // 2} packed weights but got {micro.size}")
        return cls(list(nodes), is_distance=is_distance, _micro=micro)

    @property
    def micro(self):
        """The packed int32 micro-unit triangle in STORAGE order (read-only view)."""
        view = self._micro.view()
        view.flags.writeable = False
        return view

    # -- properties (matrix.py:51-149) -------------------------------------------
    @property
    def is_distance(self):
        return self._is_distance

    @property
    def nodes(self):
        return self._nodes[:]

    @property
    def diameter(self):
        if not self.is_distance:
            raise ValueError("cannot compute diameter for similarity matrix")
        n = len(self._store)
        if n < 2:
            return 0.0
        mask = np.ones(self._micro.size, dtype=bool)
        mask[self._offsets[:-1]] = False          # drop the diagonal
        vals = self._micro[mask]
        vals = vals[vals != _UNSET]
        return round(float(vals.max()) / 1e6, 6) if vals.size else 0.0

    def _node_averages(self):
        """Per node, ``average`` of every weight it appears in while iterating the
        upper half -- the diagonal is seen as source AND target, so it counts twice
        (matrix.py:81-96); summed in the reference's order so ties break alike."""
        return {name: self._reference_average(k) for k, name in enumerate(self._nodes)}

    def _row_micro(self, k):
        """Row k (traversal order) of the full symmetric matrix, int32 micro-units."""
        order = np.fromiter((self._index[x] for x in self._nodes), dtype=np.int64, count=len(self._nodes))
        p = order[k]
        lo, hi = np.minimum(order, p), np.maximum(order, p)
        return self._micro[self._offsets[lo] + (hi - lo)]

    def _reference_average(self, k):
        """The float the reference computes for node k: ``statistics.average`` over the list
        [w(0,k) .. w(k-1,k), w(k,k), w(k,k), w(k,k+1) ..] in that order (matrix.py:84-96)."""
        row = self._row_micro(k)
        if (row == _UNSET).any():
            raise TypeError("unsupported operand type(s): matrix holds unset weights")
        vals = (row.astype(np.float64) / 1e6).tolist()
        return average(vals[:k] + [vals[k], vals[k]] + vals[k + 1:])

    def _central(self, smallest, first):
        """(name, average) of the node with the smallest / largest average weight, ties
        resolved as ``sorted(...)[0]`` (first in traversal order) or ``[-1]`` (last) would.

        No dense N x N and no per-edge Python: exact integer row sums come from ONE pass over
        the packed triangle in the library (pc_node_sums).  Two nodes whose integer sums differ
        (by >= 1e-6) cannot be reordered by the float noise of upstream's summation (< 1e-11 at
        50 000 nodes), so only nodes with EQUAL integer sums are ranked with upstream's own
        float arithmetic."""
        n = len(self._store)
        sums = np.empty(n, dtype=np.int64)
        lib = _native.load_library()
        micro = np.ascontiguousarray(self._micro)
        if lib.pc_node_sums(n, micro.ctypes.data, sums.ctypes.data):
            raise TypeError("unsupported operand type(s): matrix holds unset weights")
        order = np.fromiter((self._index[x] for x in self._nodes), dtype=np.int64, count=n)
        t_sums = sums[order]
        best = t_sums.min() if smallest else t_sums.max()
        tied = np.nonzero(t_sums == best)[0].tolist()
        ranked = [(self._reference_average(k), k) for k in tied]
        # sorted() is stable: among equal averages [0] is the first in traversal order, [-1] the last
        values = [v for v, _ in ranked]
        target = min(values) if smallest else max(values)
        ks = [k for v, k in ranked if v == target]
        k = ks[0] if first else ks[-1]
        return self._nodes[k], self._reference_average(k)

    @property
    def medoid(self):
        """matrix.py:81-104: distance -> smallest average, ``sorted(...)[0]``; similarity ->
        largest, ``sorted(...)[-1]``."""
        if len(self._store) == 0:
            raise IndexError("list index out of range")
        return self._central(smallest=True, first=True) if self.is_distance else \
            self._central(smallest=False, first=False)

    @property
    def anti_medoid(self):
        """matrix.py:106-127."""
        if len(self._store) == 0:
            raise IndexError("list index out of range")
        return self._central(smallest=False, first=False) if self.is_distance else \
            self._central(smallest=True, first=True)

    @property
    def statistics(self):
        """(mean, sample std-dev, sample skewness) of the off-diagonal edges
        (matrix.py:129-149)."""
        n = len(self)
        if n == 1:
            node = self._nodes[0]
            return self.get_weight(node, node), 0.0, 0.0
        edges = [w for s, t, w in self if s != t]
        mean = average(edges)
        var = sum([(x - mean) ** 2 for x in edges]) / (len(edges) - 1) if len(edges) > 1 else 0.0
        std = var ** 0.5
        if std == 0.0:
            return mean, std, 0.0
        skew = sum([(x - mean) ** 3 for x in edges]) / ((std ** 3) * (len(edges) - 1))
        return mean, std, skew

    # -- element access -------------------------------------------------------------
    def _slot(self, source, target):
        if source not in self:
            raise KeyError(f"node '{source}' not in matrix")
        if target not in self:
            raise KeyError(f"node '{target}' not in matrix")
        i, j = self._index[source], self._index[target]
        if i > j:
            i, j = j, i
        return int(self._offsets[i]) + (j - i)

    def get_weight(self, source, target):
        v = int(self._micro[self._slot(source, target)])
        return None if v == _UNSET else v / 1e6

    def set_weight(self, source, target, weight):
        if self._locked:
            raise AttributeError("matrix is marked as read-only")
        slot = self._slot(source, target)
        if not 0 <= weight <= 1:
            raise ValueError(f"weight {weight} not in [0.0, 1.0]")
        self._micro[slot] = int(round(round(weight, 6) * 1e6))

    def append_node(self, source, data):
        """matrix.py:169-216 -- grows the packed triangle by one node."""
        if source in self:
            raise KeyError(f"node '{source}' is already in this matrix")
        if source not in data:
            raise KeyError(f"incoming data lacks an self-edge for '{source}'")
        if self.is_distance and data[source] != 0.0:
            raise ValueError(f"nonsense value {data[source]} for self-edge on distance matrix")
        if not self.is_distance and data[source] != 1.0:
            raise ValueError(f"nonsense value {data[source]} for self-edge on similarity matrix")
        missed = set(self._store) - data.keys()
        if missed:
            raise KeyError(f"missing edge(s) for {source} vs: {missed}")
        extra = set(data.keys()) - {source} - set(self._store)
        if extra:
            raise KeyError(f"specified edges for nodes not found in matrix: {extra}")
        old_n = len(self._store)
        old = self._dense_micro()
        self._store.append(source)
        self._index[source] = old_n
        self._nodes.append(source)
        n = old_n + 1
        dense = np.full((n, n), _UNSET, dtype=np.int32)
        dense[:old_n, :old_n] = old
        self._offsets = _tri_offsets(n)
        self._micro = dense[np.triu_indices(n)].astype(np.int32)
        for target in self._nodes:
            self.set_weight(source, target, data[target])

    def extract_submatrix(self, nodes):
        for node in nodes:
            if node not in self:
                raise KeyError(f"node '{node}' not in matrix")
        idx = np.fromiter((self._index[x] for x in nodes), dtype=np.int64, count=len(nodes))
        k = len(nodes)
        if k <= 256:
            iu, ju = np.triu_indices(k)
            a, b = idx[iu], idx[ju]
            lo, hi = np.minimum(a, b), np.maximum(a, b)
            micro = self._micro[self._offsets[lo] + (hi - lo)]
            return SymMatrix(nodes, self.is_distance, _micro=np.ascontiguousarray(micro))
        # large blocks row by row: no k(k+1)/2-sized index arrays (5000 of 10 000 nodes: 1.2 s -> 0.26 s)
        micro = np.empty(k * (k + 1) // 2, dtype=np.int32)
        offsets, store = self._offsets, self._micro
        pos = 0
        for i in range(k):
            rest = idx[i:]
            lo, hi = np.minimum(idx[i], rest), np.maximum(idx[i], rest)
            micro[pos:pos + k - i] = store[offsets[lo] + (hi - lo)]
            pos += k - i
        return SymMatrix(nodes, self.is_distance, _micro=micro)

    def threshold_components(self, max_micro):
        """Connected components of the graph whose edges are the off-diagonal pairs with
        weight <= max_micro micro-units, as one int label per node of ``self.nodes``
        (traversal order); None if a weight is unset.  The packed triangle is scanned in
        row blocks (no dense N x N): this is what lets clustering.hierarchical_clustering
        hand scikit-learn one small block per component instead of the whole matrix."""
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import connected_components
        n = len(self._store)
        if (self._micro == _UNSET).any():
            return None
        rows, cols = [], []
        offsets = self._offsets
        step = max(1, (1 << 24) // max(n, 1))              # ~16 M entries per block
        for r0 in range(0, n, step):
            r1 = min(n, r0 + step)
            lo, hi = int(offsets[r0]), int(offsets[r1])
            hit = np.flatnonzero(self._micro[lo:hi] <= max_micro) + lo
            if hit.size == 0:
                continue
            i = np.searchsorted(offsets, hit, side="right") - 1
            j = i + (hit - offsets[i])
            off_diag = j != i
            rows.append(i[off_diag])
            cols.append(j[off_diag])
        if rows:
            rows, cols = np.concatenate(rows), np.concatenate(cols)
        else:
            rows = cols = np.zeros(0, dtype=np.int64)
        graph = coo_matrix((np.ones(len(rows), dtype=np.int8), (rows, cols)), shape=(n, n))
        _, storage_label = connected_components(graph, directed=False)
        return storage_label[self._traversal_order()]

    def invert(self):
        """distance <-> similarity in place: round(1 - w, 6) == (1e6 - n)/1e6."""
        if self._locked:
            raise AttributeError("matrix is marked as read-only")
        unset = self._micro == _UNSET
        if unset.any():
            raise TypeError("unsupported operand type(s) for -: 'float' and 'NoneType'")
        self._micro = (1000000 - self._micro).astype(np.int32)
        self._is_distance = not self.is_distance
        return self

    def reorder(self, nodes=None):
        if not nodes:
            nodes = [self._nodes[x] for x in _get_tree_order(self)]
        if len(nodes) != len(self):
            raise ValueError(f"need {len(self)} nodes but got {len(nodes)}")
        for node in nodes:
            if node not in self:
                raise KeyError(f"node '{node}' not in matrix")
        self._nodes = nodes

    def nearest_neighbors(self, source, threshold):
        is_distance = self.is_distance
        neighbors = []
        for target in self._nodes:
            if source == target:
                continue
            weight = self.get_weight(source, target)
            if (is_distance and weight <= threshold) or (not is_distance and weight >= threshold):
                neighbors.append(target)
        return sorted(neighbors, reverse=not is_distance,
                      key=lambda x: self.get_weight(source, x))

    def lock(self):
        self._locked = True

    def unlock(self):
        self._locked = False

    def is_locked(self):
        return self._locked

    # -- bulk views ---------------------------------------------------------------------
    def _traversal_order(self):
        """Storage index of every node in traversal order."""
        if self._nodes == self._store:
            return np.arange(len(self._store), dtype=np.int64)
        return np.fromiter((self._index[x] for x in self._nodes), dtype=np.int64, count=len(self._nodes))

    def _dense(self, order, kind):
        """n x n view of the packed triangle with rows/columns in ``order`` through
        libphamclust_b200's host-side pc_dense_from_micro (threaded gather; no
        N(N+1)/2-sized index arrays): kind 0 = float64 weights, 1 = int32 micro-units.
        Returns None when kind 0 meets an unset weight."""
        n = len(self._store)
        out = np.empty((n, n), dtype=np.float64 if kind == 0 else np.int32)
        lib = _native.load_library()
        micro = np.ascontiguousarray(self._micro, dtype=np.int32)
        rc = lib.pc_dense_from_micro(n, micro.ctypes.data, order.ctypes.data, kind, out.ctypes.data)
        return out if rc == 0 else None

    def _dense_micro(self):
        """Square int32 matrix in STORAGE order."""
        return self._dense(np.arange(len(self._store), dtype=np.int64), 1)

    def to_ndarray(self, condensed=False):
        """Square float64 matrix in TRAVERSAL order (or scipy's condensed form);
        what clustering.py:27 and _get_tree_order consume."""
        order = self._traversal_order()
        full = self._dense(order, 0)
        if full is None:                  # unset weights: the object array the reference builds
            dense = self._dense(order, 1)
            full = dense.astype(object) / 1e6
            full[dense == _UNSET] = None
            return full
        if condensed:
            from scipy.spatial.distance import squareform
            return squareform(full, force="tovector")
        return full

    def iterrows(self):
        full = self.to_ndarray()
        for k, source in enumerate(self._nodes):
            yield source, [x if x is None else float(x) for x in full[k]]

    def __iter__(self):
        """(source, target, weight) over the upper half in traversal order
        (matrix.py:349-357); one vectorised row of the packed triangle per source
        instead of one dict-style lookup per edge."""
        nodes = self._nodes[:]
        order = self._traversal_order()
        offsets = self._offsets
        for idx, source in enumerate(nodes):
            a, b = order[idx], order[idx:]
            lo, hi = np.minimum(a, b), np.maximum(a, b)
            vals = self._micro[offsets[lo] + (hi - lo)]
            weights = (vals / 1e6).tolist()
            if (vals == _UNSET).any():
                weights = [None if v == _UNSET else w for v, w in zip(vals.tolist(), weights)]
            for target, weight in zip(nodes[idx:], weights):
                yield source, target, weight

    def __contains__(self, item):
        if not isinstance(item, str):
            raise TypeError(f"type(item) should be 'str', not '{type(item)}'")
        return item in self._index

    def __getitem__(self, item):
        """Row as the reference stores it: only targets that sort AFTER (or equal
        to) ``item`` lexicographically (matrix.py:230-234, 319-323)."""
        if item not in self:
            raise KeyError(f"node '{item}' not in matrix")
        row = {}
        for target in self._store:
            if target >= item:
                w = self.get_weight(item, target)
                if w is not None:
                    row[target] = w
        return row

    def __len__(self):
        return len(self._nodes)

    def __lt__(self, other):
        if not isinstance(other, SymMatrix):
            raise TypeError(f"cannot compare SymMatrix to {type(other)}")
        return len(self) < len(other)

    def __str__(self):
        s = f"{len(self)}\n"
        for source, row in self.iterrows():
            s += f"{source:<24}\t" + "\t".join([f"{x:.6f}" for x in row]) + "\n"
        return s


# ---------------------------------------------------------------------------------------
# matrix_de_novo and friends (matrix.py:409-497)
# ---------------------------------------------------------------------------------------
def _outside_in_index_iterator(num_indices):
    """[0, n-1, 1, n-2, ...] (matrix.py:409-423).  The reference uses it to
    balance joblib batches; on the GPU the result is order independent, it is
    kept for API parity.

    >>> list(_outside_in_index_iterator(5))
    [0, 4, 1, 3, 2]
    """
    lo, hi = 0, num_indices - 1
    while lo < hi:
        yield lo
        yield hi
        lo, hi = lo + 1, hi - 1
    if lo == hi:
        yield lo


def calculate_adjacency(source, target, func, distance=True):
    """matrix.py:426-429."""
    return source.name, target.name, func(source, target, as_distance=distance)


_contexts = {}


def get_context(device=0):
    """The process-wide pc_ctx of a device (created on first use)."""
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = _native.Context(device)
    return ctx


def _metric_key(func):
    from phamclust_b200 import metrics
    key = metrics.metric_key(func)
    if key is None:
        raise TypeError(
            "matrix_de_novo runs on the GPU and only knows PhamClust's six metrics "
            "(gcs, jc, pocp, af, aai, peq from phamclust_b200.metrics); there is no CPU "
            f"fallback for arbitrary callables such as {func!r}")
    return key


def packed_matrix(table, metric, as_distance=True, ppos=False, device=0, engine=None):
    """Packed int32 micro-unit triangle for a PackedTable (single GPU)."""
    ctx = get_context(device)
    if engine is not None:
        ctx.set_engine(engine)
    ctx.load_table(table)
    if metric in ("aai", "peq"):
        if table.seq_off is None:
            raise ValueError(f"metric {metric} needs translations")
        ctx.load_sequences(table)
    return ctx.matrix(metric, as_distance=as_distance, ppos=ppos)


def matrix_de_novo(genomes, func, cpus=1, as_distance=True):
    """Drop-in for matrix.py:432-497.

    ``genomes``: list[Genome]; their order fixes the source<target orientation
    (matrix.py:480-486).  ``func``: one of the six metric functions.  ``cpus`` is
    accepted for signature parity and ignored (one GPU does the whole triangle;
    see phamclust_b200.distributed for several).  Returns a ``SymMatrix`` with
    ``is_distance == as_distance`` whose diagonal is ``1.0 - as_distance``.
    """
    if len(genomes) == 0:
        raise ValueError("need at least 1 genome to construct matrix de novo")
    key = _metric_key(func)
    n = len(genomes)
    logging.debug(f"{n} genomes -> {n * (n - 1) // 2 + n} edges including diagonal")
    table = genomes if isinstance(genomes, PackedTable) else \
        pack_genomes(genomes, with_sequences=key in ("aai", "peq"))
    from phamclust_b200 import distributed
    if distributed.is_active():
        micro = distributed.sharded_matrix(table, key, as_distance=as_distance)
    else:
        micro = packed_matrix(table, key, as_distance=as_distance)
    return SymMatrix.from_micro(table.names, micro, is_distance=bool(as_distance))


class ResidentMatrix:
    """A distance / similarity matrix whose packed triangle STAYS in HBM (the ctx's resident
    buffer, include/phamclust_b200.h "the triangle kept in HBM").  It offers what the pipeline's
    first consumer needs -- thresholded clustering (scripts/phamclust.py:300-317) through
    ``clustering.hierarchical_clustering`` -- without shipping N(N+1)/2 values to the host:
    ``threshold_components`` asks the device for the pairs under the threshold (an edge list),
    ``extract_submatrix`` gathers just a component's entries into an ordinary host ``SymMatrix``.
    ``to_symmatrix()`` fetches everything after all.  Valid until the next matrix is made resident
    or another table is loaded on that context (checked)."""

    def __init__(self, names, ctx, is_distance, generation):
        self._nodes = list(names)
        self._index = {name: k for k, name in enumerate(self._nodes)}
        self._ctx, self._is_distance, self._generation = ctx, bool(is_distance), generation
        self._offsets = _tri_offsets(len(self._nodes))
        self.d2h_bytes = 0                       # what the queries have shipped so far

    def _check(self):
        if getattr(self._ctx, "resident_generation", None) != self._generation or self._ctx.n != len(self._nodes):
            raise RuntimeError("the resident matrix of this context has been replaced")

    is_distance = property(lambda self: self._is_distance)
    nodes = property(lambda self: self._nodes[:])

    def __len__(self):
        return len(self._nodes)

    def __contains__(self, item):
        return item in self._index

    def threshold_components(self, max_micro):
        """One component label per node (the graph of off-diagonal pairs with weight <=
        ``max_micro`` micro-units), like ``SymMatrix.threshold_components`` -- the scan runs on
        the device, only the edge list crosses PCIe."""
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import connected_components
        self._check()
        n = len(self._nodes)
        edges = self._ctx.resident_threshold_edges(int(max_micro))
        self.d2h_bytes += edges.nbytes
        graph = coo_matrix((np.ones(len(edges), dtype=np.int8), (edges[:, 0], edges[:, 1])), shape=(n, n))
        _, label = connected_components(graph, directed=False)
        return label

    def extract_submatrix(self, nodes):
        self._check()
        for node in nodes:
            if node not in self._index:
                raise KeyError(f"node '{node}' not in matrix")
        idx = np.fromiter((self._index[x] for x in nodes), dtype=np.int64, count=len(nodes))
        iu, ju = np.triu_indices(len(nodes))
        a, b = idx[iu], idx[ju]
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        micro = self._ctx.resident_gather(self._offsets[lo] + (hi - lo))
        self.d2h_bytes += micro.nbytes
        return SymMatrix(list(nodes), self._is_distance, _micro=micro)

    def to_symmatrix(self):
        self._check()
        micro = self._ctx.resident_fetch()
        self.d2h_bytes += micro.nbytes
        return SymMatrix.from_micro(self._nodes, micro, self._is_distance)

    def to_ndarray(self, condensed=False):
        return self.to_symmatrix().to_ndarray(condensed)


def matrix_de_novo_resident(genomes, func, cpus=1, as_distance=True, device=0):
    """``matrix_de_novo`` that leaves the result in HBM and returns a ``ResidentMatrix``
    (single GPU).  Same arguments and errors as ``matrix_de_novo``."""
    if len(genomes) == 0:
        raise ValueError("need at least 1 genome to construct matrix de novo")
    key = _metric_key(func)
    table = genomes if isinstance(genomes, PackedTable) else \
        pack_genomes(genomes, with_sequences=key in ("aai", "peq"))
    ctx = get_context(device)
    ctx.load_table(table)
    if key in ("aai", "peq"):
        if table.seq_off is None:
            raise ValueError(f"metric {key} needs translations")
        ctx.load_sequences(table)
    generation = ctx.matrix_resident(key, as_distance=as_distance)
    return ResidentMatrix(table.names, ctx, as_distance, generation)


# ---------------------------------------------------------------------------------------
# file formats (matrix.py:500-670): adjacency, squareform, lower triangle; "%.6f"
# ---------------------------------------------------------------------------------------
def read_adjacency(filepath):
    with open(filepath, "r") as handle:
        for edge in handle:
            edge = edge.rstrip().split("\t")
            yield edge[0], edge[1], float(edge[2])


def matrix_from_adjacency(filepath):
    names, diagonal = dict(), set()
    for source, target, weight in read_adjacency(filepath):
        if source == target:
            diagonal.add(weight)
        if target in names:
            break
        names[target] = None
    names = list(names.keys())
    if len(diagonal) > 1:
        raise ValueError("values on matrix diagonal should be identical")
    if sum(diagonal) == 0.0:
        is_distance = True
    elif sum(diagonal) == 1.0:
        is_distance = False
    else:
        raise ValueError("values on matrix diagonal can only be 0.0 or 1.0")
    matrix = SymMatrix(names, is_distance=is_distance)
    for source, target, weight in read_adjacency(filepath):
        matrix.set_weight(source, target, weight)
    return matrix


def _fmt6(micro):
    """'%.6f' of micro-units without going through floats."""
    micro = int(micro)
    return f"{micro // 1000000}.{micro % 1000000:06d}"


def _native_write(matrix, filepath, mode):
    """Write through libphamclust_b200's host-side formatter (pc_write_matrix):
    integer micro-units -> '%.6f' text, no per-value Python.  Returns False when
    the matrix is not array-backed or holds unset weights (caller falls back to
    the plain Python writer, which then fails exactly like the reference)."""
    if not isinstance(matrix, SymMatrix) or (matrix._micro == _UNSET).any():
        return False
    import ctypes
    lib = _native.load_library()
    names = [x.encode() for x in matrix._store]
    off = np.zeros(len(names) + 1, dtype=np.int64)
    np.cumsum([len(x) for x in names], out=off[1:])
    blob = b"".join(names)
    order = np.fromiter((matrix._index[x] for x in matrix._nodes), dtype=np.int64, count=len(matrix._nodes))
    micro = np.ascontiguousarray(matrix._micro, dtype=np.int32)
    rc = lib.pc_write_matrix(str(filepath).encode(), len(names), micro.ctypes.data, order.ctypes.data,
                             ctypes.c_char_p(blob), off.ctypes.data, mode)
    if rc:
        raise ValueError(lib.pc_io_last_error().decode())
    return True


def matrix_to_adjacency(matrix, filepath, skip_zero=False):
    if _native_write(matrix, filepath, 3 if skip_zero else 2):
        return filepath
    with open(filepath, "w") as handle:
        for source, target, weight in matrix:
            if skip_zero and not weight:
                continue
            handle.write(f"{source}\t{target}\t{weight:.6f}\n")
    return filepath


def read_squareform(filepath):
    with open(filepath, "r") as handle:
        next(handle)                               # first line only has the size
        for row in handle:
            row = row.rstrip().split("\t")
            yield row[0], [float(x) for x in row[1:]]


def matrix_from_squareform(filepath):
    """Squareform / lower-triangle reader (matrix.py:564-603): the whole file is
    parsed by pc_parse_squareform into the packed triangle; diagonal checks as
    upstream."""
    with open(filepath, "rb") as handle:
        data = handle.read()
    first = data.split(b"\n", 1)[0].split(b"\t")[0]
    n = int(first)
    lib = _native.load_library()
    micro = np.empty(n * (n + 1) // 2, dtype=np.int32)
    nb = np.empty(max(n, 1), dtype=np.int64)
    ne = np.empty(max(n, 1), dtype=np.int64)
    rc = lib.pc_parse_squareform(data, len(data), n, micro.ctypes.data, nb.ctypes.data, ne.ctypes.data)
    if rc:
        raise ValueError(lib.pc_io_last_error().decode())
    names = [data[nb[i]:ne[i]].decode() for i in range(n)]
    diag = micro[_tri_offsets(n)[:-1]] if n else micro
    if len(set(diag.tolist())) > 1:
        raise ValueError("values on matrix diagonal should be identical")
    if n and int(diag[0]) == 0:
        is_distance = True
    elif n and int(diag[0]) == 1000000:
        is_distance = False
    else:
        raise ValueError("values on matrix diagonal can only be 0.0 or 1.0")
    return SymMatrix(names, is_distance=is_distance, _micro=micro)


def _matrix_from_squareform_python(filepath):
    """The reference's two-pass reader, kept for cross-checking the native parser."""
    names, diagonal = [], set()
    for i, (target, row) in enumerate(read_squareform(filepath)):
        names.append(target)
        diagonal.add(row[i])
    if len(diagonal) > 1:
        raise ValueError("values on matrix diagonal should be identical")
    if sum(diagonal) == 0.0:
        is_distance = True
    elif sum(diagonal) == len(diagonal):
        is_distance = False
    else:
        raise ValueError("values on matrix diagonal can only be 0.0 or 1.0")
    matrix = SymMatrix(names, is_distance=is_distance)
    for i, (target, row) in enumerate(read_squareform(filepath)):
        for source, weight in zip(names[:i + 1], row[:i + 1]):
            matrix.set_weight(source, target, weight)
    return matrix


def matrix_to_squareform(matrix, filepath, lower_triangle=False):
    """Squareform / lower-triangle writer; the lower-triangle form is the
    pipeline's on-disk cache (scripts/phamclust.py:254-262) and must stay
    byte-identical: '%.6f', tab separated."""
    if _native_write(matrix, filepath, 1 if lower_triangle else 0):
        return filepath
    with open(filepath, "w") as handle:
        header = f"{len(matrix)}"
        if not lower_triangle:
            header += "\t" + "\t".join(matrix.nodes)
        handle.write(f"{header}\n")
        for i, (source, row) in enumerate(matrix.iterrows()):
            if lower_triangle:
                row = row[:i + 1]
            handle.write(f"{source}\t" + "\t".join([f"{x:.6f}" for x in row]) + "\n")
    return filepath


def _get_tree_order(matrix):
    """Leaf order of a single-linkage dendrogram (matrix.py:673-686): what
    ``dendrogram(z, no_plot=True, get_leaves=True, count_sort="descending")["leaves"]`` returns --
    at every merge the child holding more leaves comes first, the second child on ties -- walked
    with an explicit stack: scipy's dendrogram recurses once per level of the tree, and a
    single-linkage tree of a few thousand nodes is often a chain deeper than the interpreter's
    recursion limit (upstream fails there with RecursionError)."""
    from scipy.cluster.hierarchy import linkage
    if not matrix.is_distance:
        matrix = matrix.extract_submatrix(matrix.nodes)
        matrix.invert()
    n = len(matrix)
    if n < 2:
        return list(range(n))
    z = linkage(matrix.to_ndarray(condensed=True), method="single")
    left = z[:, 0].astype(np.int64)
    right = z[:, 1].astype(np.int64)

    def count(node):
        return 1 if node < n else int(z[node - n, 3])

    leaves = []
    stack = [2 * n - 2]
    while stack:
        node = stack.pop()
        if node < n:
            leaves.append(int(node))
            continue
        a, b = int(left[node - n]), int(right[node - n])
        first, second = (a, b) if count(a) > count(b) else (b, a)
        stack.append(second)
        stack.append(first)
    return leaves
