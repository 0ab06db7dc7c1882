"""Host hierarchical clustering hand-off (scikit-learn's AgglomerativeClustering on the
precomputed distance matrix), mirroring /root/reference/src/phamclust/clustering.py:4-51.
Not accelerated -- it is the consumer the device-built ``SymMatrix`` must satisfy.

Large matrices: the reference exports a dense float64 N x N
(clustering.py:27), 20 GB at 50 000 genomes before scikit-learn copies it.  With a
distance threshold, single / average / complete linkage never merge two clusters whose
every cross pair is at or above the threshold (the linkage distance is a min / mean / max
of those pairs), so the flat clustering decomposes EXACTLY over the connected components of
the graph {d < eps}: components come from one thresholded scan of the packed int32
triangle, and only each component's own small dense block goes through the unchanged
scikit-learn call.  Partitions and their order are those of the dense path
(tests/test_host_api.py compares them).  Caveat, shared with the reference: when several
merges tie EXACTLY, average / complete linkage have more than one valid flat clustering and
scipy's nearest-neighbour chain picks by visiting order, so on tie-heavy inputs (distances on
a coarse grid) the component-wise and the dense run may each return a different valid one;
single linkage is unique.  Matrices below ``DECOMPOSE_MIN_NODES`` nodes always take the
reference's dense path.
"""
from sklearn.cluster import AgglomerativeClustering

# matrices with at least this many nodes take the component-wise path (tests lower it)
DECOMPOSE_MIN_NODES = 4096
_DECOMPOSABLE = ("single", "average", "complete")


def hierarchical_clustering(matrix, linkage, eps=None, n_clusters=None):
    """Cluster the nodes of a DISTANCE matrix; returns the cluster
    sub-matrices, largest first.  ``eps`` and ``n_clusters`` are exclusive."""
    if len(matrix) == 1:
        return [matrix]
    if not matrix.is_distance:
        raise ValueError("matrix must be a distance matrix")
    if (eps is not None and not n_clusters and eps > 0 and linkage in _DECOMPOSABLE
            and len(matrix) >= DECOMPOSE_MIN_NODES and hasattr(matrix, "threshold_components")):
        labels = _labels_by_components(matrix, linkage, eps)
        if labels is not None:
            return _clusters_from_labels(matrix, labels)
    dist_mat = matrix.to_ndarray()
    if eps is None and not n_clusters:
        raise ValueError("need either threshold or n_clusters to proceed")
    if eps and n_clusters:
        raise ValueError("threshold and n_clusters are mutually exclusive")
    model = AgglomerativeClustering(metric="precomputed", linkage=linkage,
                                    distance_threshold=eps, n_clusters=n_clusters)
    return _clusters_from_labels(matrix, model.fit_predict(dist_mat))


def _clusters_from_labels(matrix, labels):
    """clustering.py:38-51: groups keyed in first-occurrence order of their label while
    walking ``matrix.nodes``, node names in a set, sub-matrices sorted largest first
    (stable: equal sizes keep that order)."""
    groups = dict()
    for name, label in zip(matrix.nodes, labels):
        groups.setdefault(label, set()).add(name)
    clusters = [matrix.extract_submatrix(list(names)) for names in groups.values()]
    return sorted(clusters, reverse=True)


def _labels_by_components(matrix, linkage, eps):
    """One hashable label per node of ``matrix.nodes``: (component, label inside the
    component).  None when the matrix holds unset weights (the dense path then fails
    exactly like the reference)."""
    # one micro-unit of slack: a coarser graph can only merge components, which keeps the
    # decomposition exact, while a mean of values == eps may round to just below eps
    comp = matrix.threshold_components(int(round(eps * 1e6)) + 1)
    if comp is None:
        return None
    nodes = matrix.nodes
    members = dict()
    for position, c in enumerate(comp.tolist()):
        members.setdefault(c, []).append(position)
    labels = [None] * len(nodes)
    for c, positions in members.items():
        if len(positions) == 1:
            labels[positions[0]] = (c, 0)
            continue
        block = matrix.extract_submatrix([nodes[p] for p in positions])     # traversal order kept
        model = AgglomerativeClustering(metric="precomputed", linkage=linkage, distance_threshold=eps,
                                        n_clusters=None)
        for p, local in zip(positions, model.fit_predict(block.to_ndarray()).tolist()):
            labels[p] = (c, local)
    return labels
