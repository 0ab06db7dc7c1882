"""Host hierarchical clustering hand-off (unchanged algorithm: scikit-learn's
AgglomerativeClustering on the precomputed distance matrix), mirroring
/root/reference/src/phamclust/clustering.py:4-51.  Not accelerated -- it is the
consumer the device-built ``SymMatrix`` must satisfy."""
from sklearn.cluster import AgglomerativeClustering


def hierarchical_clustering(matrix, linkage, eps=None, n_clusters=None):
    """Cluster the nodes of a DISTANCE matrix; returns the cluster
    sub-matrices, largest first.  ``eps`` and ``n_clusters`` are exclusive."""
    if len(matrix) == 1:
        return [matrix]
    if not matrix.is_distance:
        raise ValueError("matrix must be a distance matrix")
    dist_mat = matrix.to_ndarray()
    if eps is None and not n_clusters:
        raise ValueError("need either threshold or n_clusters to proceed")
    if eps and n_clusters:
        raise ValueError("threshold and n_clusters are mutually exclusive")
    model = AgglomerativeClustering(metric="precomputed", linkage=linkage,
                                    distance_threshold=eps, n_clusters=n_clusters)
    groups = dict()
    for name, label in zip(matrix.nodes, model.fit_predict(dist_mat)):
        groups.setdefault(label, set()).add(name)
    clusters = [matrix.extract_submatrix(list(names)) for names in groups.values()]
    return sorted(clusters, reverse=True)
