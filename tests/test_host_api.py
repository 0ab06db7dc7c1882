"""Host-side mirror of the reference interface: Genome, parsers, packing,
SymMatrix, file formats, and the C-ABI library's symbol table (no GPU calls)."""
import ctypes
import pathlib
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from oracle import testdata
from phamclust_b200 import (Genome, SymMatrix, load_genomes_from_fasta_dir,
                            load_genomes_from_tsv, matrix_from_adjacency,
                            matrix_from_squareform, matrix_to_adjacency, matrix_to_squareform,
                            pack_genomes)
from phamclust_b200 import _native, distributed, metrics
from phamclust_b200.matrix import _outside_in_index_iterator, matrix_de_novo
from phamclust_b200.packing import unpack_genomes


# ----------------------------------------------------------------------------- library
def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "phamclust_b200.h").read_text()
    declared = set(re.findall(r"^\s*(?:int64_t|int|void|const char \*)\s*\**(pc_\w+)\s*\(", header, re.M))
    assert declared == set(_native.EXPORTS)
    lib = ctypes.CDLL(str(_native.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pc_version() >= 100


def test_product_path_never_imports_the_oracle():
    for path in (ROOT / "phamclust_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path


# ----------------------------------------------------------------------------- genome
def test_genome_api():
    g, h = Genome("g"), Genome("h")
    for p, t in (("p1", "MKT"), ("p1", "MKTAA"), ("p2", "MKT")):
        g.add(p, t)
    h.add("p2", "MKTW")
    h.add("p3")
    assert len(g) == 3 and len(g.phams) == 2 and g["p1"] == ["MKT", "MKTAA"]
    assert g & h == {"p2"} and g | h == {"p1", "p2", "p3"} and g - h == {"p1"}
    assert g ^ h == {"p1", "p3"} and g.intersection(h) == {"p2"} and h["p3"] == ["M"]
    assert "p1" in g and "zz" not in g and h < g
    assert str(g) == ">name=g|pham=p1|n=1\nMKT\n>name=g|pham=p1|n=2\nMKTAA\n>name=g|pham=p2|n=1\nMKT\n"
    with pytest.raises(TypeError):
        g.add(1, "M")
    with pytest.raises(TypeError):
        g.add("p", 3)
    with pytest.raises(KeyError):
        g["nope"]
    with pytest.raises(TypeError):
        g & "h"
    assert g.pop("nope") is None and g.pop("p2") == ["MKT"]


def test_genome_matches_reference_class(reference):
    rows = testdata.genome_rows(seed=5, n_genomes=4, n_phams=6)
    mine = testdata.rows_to_genomes(rows, cls=Genome)
    theirs = testdata.rows_to_genomes(rows, cls=reference.genome.Genome)
    for a, b in zip(mine, theirs):
        assert str(a) == str(b) and len(a) == len(b) and a.phams == b.phams
    assert (mine[0] & mine[1]) == (theirs[0] & theirs[1])


def test_parsers(tmp_path):
    tsv = tmp_path / "in.tsv"
    tsv.write_text("a\tp1\tMKT\na\tp1\tMKTAA\nb\tp2\nb\tp3\t\na\tp2\tMKT\n")
    genomes = load_genomes_from_tsv(tsv)
    assert [g.name for g in genomes] == ["a", "b"]
    assert genomes[0].phams == {"p1": ["MKT", "MKTAA"], "p2": ["MKT"]}
    assert genomes[1].phams == {"p2": ["M"], "p3": ["M"]}       # rstrip eats the empty column
    bad = tmp_path / "bad.tsv"
    bad.write_text("a\tb\tc\td\n")
    with pytest.raises(ValueError):
        load_genomes_from_tsv(bad)
    d = tmp_path / "fa"
    d.mkdir()
    genomes[0].save(d / "a.fasta")
    (d / "b.faa").write_text(">name=b|pham=p9|n=1\nMK\nTW\n")
    (d / "notes.txt").write_text("skip me")
    loaded = {g.name: g for g in load_genomes_from_fasta_dir(d)}
    assert loaded["a"].phams == genomes[0].phams and loaded["b"].phams == {"p9": ["MKTW"]}
    (d / "a.fa").write_text(">name=a|pham=p1|n=1\nM\n")
    with pytest.raises(ValueError):
        load_genomes_from_fasta_dir(d)


def test_parsers_match_reference(reference, tmp_path):
    import importlib
    rows = testdata.genome_rows(seed=8, n_genomes=5, n_phams=7)
    tsv = tmp_path / "in.tsv"
    tsv.write_text("".join(f"{a}\t{b}\t{c}\n" for a, b, c in rows))
    mine = load_genomes_from_tsv(tsv)
    theirs = testdata.rows_to_genomes(rows, cls=reference.genome.Genome)
    assert [str(g) for g in mine] == [str(g) for g in theirs]


# ----------------------------------------------------------------------------- packing
def test_pack_round_trip():
    rows = testdata.genome_rows(seed=3, n_genomes=7, n_phams=9, paralog_p=0.5)
    genomes = testdata.rows_to_genomes(rows, cls=Genome)
    t = pack_genomes(genomes)
    assert t.n_genomes == 7 and t.n_genes == sum(len(g) for g in genomes)
    assert int(t.len_sum.sum()) == sum(len(s) for g in genomes for v in g.phams.values() for s in v)
    back = unpack_genomes(t)
    assert [g.phams for g in back] == [g.phams for g in genomes]
    sub = t.subset(3)
    assert [g.phams for g in unpack_genomes(sub)] == [g.phams for g in genomes[:3]]


def test_direct_tsv_loader_equals_genome_path(tmp_path):
    import random
    from phamclust_b200.packing import table_from_tsv
    rows = testdata.genome_rows(seed=3, n_genomes=9, n_phams=9, paralog_p=0.5)
    random.Random(1).shuffle(rows)           # paralogs and genomes interleaved in the file
    tsv = tmp_path / "x.tsv"
    tsv.write_text("".join(f"{a}\t{b}\t{c}\n" for a, b, c in rows) + "zz\tq1\nzz\tq2\t\nzz\tq1\tMKT\n")
    a, b = pack_genomes(load_genomes_from_tsv(tsv)), table_from_tsv(tsv)
    assert a.names == b.names and a.phams == b.phams
    for f in ("genome_ptr", "pham_id", "copies", "len_sum", "seq_off", "residues"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    bad = tmp_path / "bad.tsv"
    bad.write_text("a\tb\tc\td\n")
    with pytest.raises(ValueError):
        table_from_tsv(bad)


def test_synthetic_tables_are_deterministic_and_consistent():
    from phamclust_b200 import synthetic
    a = synthetic.phage_table(100, seed=42)
    b = synthetic.phage_table(100, seed=42)
    assert np.array_equal(a.residues, b.residues) and np.array_equal(a.pham_id, b.pham_id)
    lens = np.diff(a.seq_off)
    entry = np.repeat(np.arange(len(a.pham_id)), a.copies)
    assert np.array_equal(np.bincount(entry, weights=lens, minlength=len(a.pham_id)).astype(np.uint32),
                          a.len_sum)
    assert lens.min() >= 10 and lens.max() <= 1600 and a.copies.max() <= 3
    s = synthetic.sparse_table(300, 3000, seed=1, clusters=10)
    for g in range(300):
        ids = s.pham_id[s.genome_ptr[g]:s.genome_ptr[g + 1]]
        assert len(set(ids.tolist())) == len(ids) and 20 <= len(ids) <= 300


# ----------------------------------------------------------------------------- SymMatrix
def _filled(cls, names, is_distance, seed=0):
    rng = np.random.default_rng(seed)
    m = cls(list(names), is_distance=is_distance)
    for i, a in enumerate(names):
        m.set_weight(a, a, 0.0 if is_distance else 1.0)
        for b in names[i + 1:]:
            m.set_weight(a, b, float(rng.integers(0, 1000001)) / 1e6 + 4e-7 * (rng.random() < 0.3))
    return m


def test_symmatrix_behaves_like_the_reference(reference):
    names = ["zeta", "alpha", "mu", "beta", "omega", "chi"]
    for is_distance in (True, False):
        mine = _filled(SymMatrix, names, is_distance, seed=4)
        ref = _filled(reference.matrix.SymMatrix, names, is_distance, seed=4)
        assert list(mine) == list(ref)
        assert [r for r in mine.iterrows()] == [r for r in ref.iterrows()]
        assert np.array_equal(mine.to_ndarray(), ref.to_ndarray())
        assert mine.medoid == ref.medoid and mine.anti_medoid == ref.anti_medoid
        assert mine.statistics == pytest.approx(ref.statistics)
        if is_distance:
            assert mine.diameter == ref.diameter
            assert np.array_equal(mine.to_ndarray(condensed=True), ref.to_ndarray(condensed=True))
        assert mine.nearest_neighbors("mu", 0.5) == ref.nearest_neighbors("mu", 0.5)
        assert mine["beta"] == ref["beta"] and str(mine) == str(ref)
        sub_m, sub_r = mine.extract_submatrix(["omega", "alpha", "mu"]), \
            ref.extract_submatrix(["omega", "alpha", "mu"])
        assert list(sub_m) == list(sub_r)
        mine.reorder()
        ref.reorder()
        assert mine.nodes == ref.nodes and list(mine) == list(ref)
        mine.invert()
        ref.invert()
        assert list(mine) == list(ref) and mine.is_distance == ref.is_distance
        mine.reorder(names[::-1])
        ref.reorder(names[::-1])
        assert np.array_equal(mine.to_ndarray(), ref.to_ndarray())
        data = {n: 0.25 for n in names}
        data["new"] = 0.0 if mine.is_distance else 1.0
        mine.append_node("new", dict(data))
        ref.append_node("new", dict(data))
        assert list(mine) == list(ref)


def test_symmatrix_errors_and_rounding():
    m = SymMatrix(["a", "b"], is_distance=True)
    assert m.get_weight("a", "b") is None
    m.set_weight("b", "a", 0.1234565)
    assert m.get_weight("a", "b") == round(0.1234565, 6)
    with pytest.raises(ValueError):
        m.set_weight("a", "b", 1.5)
    with pytest.raises(KeyError):
        m.get_weight("a", "zz")
    with pytest.raises(TypeError):
        3 in m
    m.lock()
    with pytest.raises(AttributeError):
        m.set_weight("a", "b", 0.5)
    m.unlock()
    assert not m.is_locked() and len(m) == 2 and m < SymMatrix(["x", "y", "z"])
    s = SymMatrix(["a"], is_distance=False)
    with pytest.raises(ValueError):
        s.diameter


def test_from_micro_and_file_formats(tmp_path):
    names = ["n3", "n1", "n2"]
    micro = np.array([0, 250000, 1000000, 0, 428571, 0], dtype=np.int32)
    m = SymMatrix.from_micro(names, micro, is_distance=True)
    assert m.get_weight("n2", "n1") == 0.428571 and m.get_weight("n3", "n2") == 1.0
    lower = matrix_to_squareform(m, tmp_path / "lower.tsv", lower_triangle=True)
    assert pathlib.Path(lower).read_text() == \
        "3\nn3\t0.000000\nn1\t0.250000\t0.000000\nn2\t1.000000\t0.428571\t0.000000\n"
    assert list(matrix_from_squareform(lower)) == list(m)
    square = matrix_to_squareform(m, tmp_path / "square.tsv")
    assert list(matrix_from_squareform(square)) == list(m)
    adj = matrix_to_adjacency(m, tmp_path / "adj.tsv")
    assert list(matrix_from_adjacency(adj)) == list(m)
    with pytest.raises(ValueError):
        SymMatrix.from_micro(names, micro[:-1], True)


def test_file_formats_match_reference_bytes(reference, tmp_path):
    names = ["zeta", "alpha", "mu", "beta"]
    mine = _filled(SymMatrix, names, True, seed=9)
    ref = _filled(reference.matrix.SymMatrix, names, True, seed=9)
    for lower in (True, False):
        a = matrix_to_squareform(mine, tmp_path / "a.tsv", lower_triangle=lower)
        b = reference.matrix.matrix_to_squareform(ref, tmp_path / "b.tsv", lower_triangle=lower)
        assert pathlib.Path(a).read_bytes() == pathlib.Path(b).read_bytes()
    a = matrix_to_adjacency(mine, tmp_path / "a.adj", skip_zero=True)
    b = reference.matrix.matrix_to_adjacency(ref, tmp_path / "b.adj", skip_zero=True)
    assert pathlib.Path(a).read_bytes() == pathlib.Path(b).read_bytes()


def test_native_writers_and_reader_match_the_python_reference_formats(reference, tmp_path):
    """pc_write_matrix / pc_parse_squareform against the verbatim reference's own
    writers and readers, after a reorder (traversal != storage order)."""
    from phamclust_b200.matrix import _matrix_from_squareform_python
    names = [f"node {k} x" if k % 3 else f"n{k}" for k in range(23)]
    mine = _filled(SymMatrix, names, True, seed=5)
    ref = _filled(reference.matrix.SymMatrix, names, True, seed=5)
    order = names[::2] + names[1::2]
    mine.reorder(list(order))
    ref.reorder(list(order))
    for lower in (True, False):
        a = matrix_to_squareform(mine, tmp_path / "a.tsv", lower_triangle=lower)
        b = reference.matrix.matrix_to_squareform(ref, tmp_path / "b.tsv", lower_triangle=lower)
        assert pathlib.Path(a).read_bytes() == pathlib.Path(b).read_bytes()
        back = matrix_from_squareform(a)
        theirs = reference.matrix.matrix_from_squareform(b)
        assert list(back) == list(theirs) and back.is_distance == theirs.is_distance
        assert list(_matrix_from_squareform_python(a)) == list(back)
    for skip in (True, False):
        a = matrix_to_adjacency(mine, tmp_path / "a.adj", skip_zero=skip)
        b = reference.matrix.matrix_to_adjacency(ref, tmp_path / "b.adj", skip_zero=skip)
        assert pathlib.Path(a).read_bytes() == pathlib.Path(b).read_bytes()
    sim = _filled(SymMatrix, names[:5], False, seed=6)
    path = matrix_to_squareform(sim, tmp_path / "s.tsv", lower_triangle=True)
    assert not matrix_from_squareform(path).is_distance
    # values with more than six decimals are rounded exactly like set_weight does
    (tmp_path / "odd.tsv").write_text("2\nx\t0.0\ny\t0.1234565\t0.0\n")
    assert matrix_from_squareform(tmp_path / "odd.tsv").get_weight("x", "y") == round(0.1234565, 6)
    (tmp_path / "bad.tsv").write_text("2\nx\t0.0\ny\t1.5\t0.0\n")
    with pytest.raises(ValueError):
        matrix_from_squareform(tmp_path / "bad.tsv")


def test_clustering_on_golden_matrix_gives_golden_partitions():
    """The unchanged host clustering on a SymMatrix built from the reference's
    own weights must reproduce the reference's partitions."""
    from phamclust_b200.clustering import hierarchical_clustering
    for dataset, golden in load_golden("matrices.json").items():
        for key in ("gcs", "peq"):
            m = SymMatrix.from_micro(golden["names"], np.array(golden["distance"][key], np.int32), True)
            for spec, expect in golden["clusters"][key].items():
                linkage, eps = spec.split(":")
                got = hierarchical_clustering(m, linkage, eps=float(eps))
                assert sorted(sorted(c.nodes) for c in got) == expect
                sizes = [len(c) for c in got]
                assert sizes == sorted(sizes, reverse=True)


# ----------------------------------------------------------------------------- matrix_de_novo (host logic)
def test_outside_in_index_iterator():
    assert list(_outside_in_index_iterator(5)) == [0, 4, 1, 3, 2]
    assert list(_outside_in_index_iterator(4)) == [0, 3, 1, 2]
    assert list(_outside_in_index_iterator(0)) == [] and list(_outside_in_index_iterator(1)) == [0]


def test_matrix_de_novo_argument_errors():
    with pytest.raises(ValueError):
        matrix_de_novo([], metrics.gene_content_similarity, 1)
    with pytest.raises(TypeError):
        matrix_de_novo([Genome("a")], lambda s, t, as_distance=False: 0.0, 1)
    assert metrics.metric_key(metrics.METRICS["peq"]) == "peq"
    assert metrics.metric_key("jc") == "jc" and metrics.metric_key(print) is None


def test_no_gpu_means_loud_failure_not_fallback():
    ok = True
    try:
        ctx = _native.Context(0)
        ctx.close()
    except RuntimeError as exc:
        ok = False
        assert "no CPU fallback" in str(exc) or "sm_" in str(exc)
    if not ok:
        with pytest.raises(RuntimeError):
            metrics.gene_content_similarity(Genome("a"), Genome("b"))


# ----------------------------------------------------------------------------- sharding arithmetic
def test_row_partition_balances_pairs():
    for n in (1, 2, 7, 100, 2001):
        for world in (1, 2, 3, 8):
            b = distributed.row_partition(n, world)
            assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
            sizes = [distributed.tri_offset(n, b[r + 1]) - distributed.tri_offset(n, b[r])
                     for r in range(world)]
            assert sum(sizes) == n * (n + 1) // 2
            if n >= 100:
                assert max(sizes) - min(sizes) <= 2 * n


def test_dense_views_come_from_the_native_gather():
    """SymMatrix.to_ndarray / _dense_micro go through pc_dense_from_micro (threaded tiled
    gather in libphamclust_b200, host only): storage order, permuted traversal order, sizes
    that are not multiples of the tile, and the object-array fallback for unset weights."""
    import numpy as np
    from phamclust_b200.matrix import SymMatrix
    rng = np.random.default_rng(12)
    for n in (1, 2, 63, 64, 65, 700):
        tri = rng.integers(0, 1000001, size=n * (n + 1) // 2, dtype=np.int32)
        names = [f"n{k:04d}" for k in range(n)]
        m = SymMatrix.from_micro(names, tri, is_distance=False)
        iu = np.triu_indices(n)
        want = np.empty((n, n), dtype=np.int32)
        want[iu] = tri
        want.T[iu] = tri
        assert np.array_equal(m._dense_micro(), want)
        assert np.array_equal(m.to_ndarray(), want / 1e6)
        perm = rng.permutation(n)
        m._nodes = [names[k] for k in perm]
        got = m.to_ndarray()
        assert got.dtype == np.float64 and np.array_equal(got, (want / 1e6)[np.ix_(perm, perm)])
    holes = SymMatrix(["a", "b", "c"], is_distance=True)
    holes.set_weight("a", "b", 0.25)
    full = holes.to_ndarray()
    assert full.dtype == object and full[0][1] == 0.25 and full[1][0] == 0.25 and full[2][2] is None


def test_component_wise_clustering_equals_the_dense_path():
    """hierarchical_clustering on matrices too large for a dense N x N: connected components
    of {d < eps} from the packed triangle, scikit-learn per component.  Same partitions in the
    same order as the reference's dense path for all three linkages (continuous distances: no
    exact ties), permuted traversal order included; single linkage also on a coarse grid."""
    import numpy as np
    from phamclust_b200 import clustering
    from phamclust_b200.matrix import SymMatrix

    def random_matrix(rng, n, k, grid):
        g = rng.integers(0, k, size=n)
        d = np.where(g[:, None] == g[None, :], rng.random((n, n)) * 0.5, 0.4 + rng.random((n, n)) * 0.6)
        d = (d + d.T) / 2
        if grid:
            d = np.round(d * grid) / grid
        np.fill_diagonal(d, 0.0)
        micro = np.rint(d[np.triu_indices(n)] * 1e6).astype(np.int32)
        return SymMatrix.from_micro([f"n{i:04d}" for i in range(n)], micro, True)

    def run(matrix, linkage, eps, min_nodes):
        saved = clustering.DECOMPOSE_MIN_NODES
        clustering.DECOMPOSE_MIN_NODES = min_nodes
        try:
            return [frozenset(c.nodes) for c in clustering.hierarchical_clustering(matrix, linkage, eps=eps)]
        finally:
            clustering.DECOMPOSE_MIN_NODES = saved

    rng = np.random.default_rng(5)
    for trial in range(30):
        n, k = int(rng.integers(5, 120)), int(rng.integers(1, 8))
        m = random_matrix(rng, n, k, 0)
        if trial % 3 == 0:
            m.reorder([m.nodes[i] for i in rng.permutation(n)])
        for linkage in ("single", "average", "complete"):
            for eps in (0.25, 0.4, 0.75):
                assert run(m, linkage, eps, 2) == run(m, linkage, eps, 10 ** 9), (trial, linkage, eps)
        coarse = random_matrix(rng, n, k, 8)
        assert run(coarse, "single", 0.4, 2) == run(coarse, "single", 0.4, 10 ** 9)
    # the scan walks the triangle in row blocks: a matrix spanning several of them
    big = random_matrix(rng, 4500, 90, 0)
    assert run(big, "complete", 0.25, 4096) == run(big, "complete", 0.25, 10 ** 9)
    # unset weights: no component labels, the dense path takes over (and fails like upstream)
    holes = SymMatrix(["a", "b", "c"], is_distance=True)
    assert holes.threshold_components(250001) is None


def test_c_packer_equals_python_packer_and_rejects_wide_characters():
    """pack_genomes walks the Genome dicts once in C (csrc/pc_pypack.c, include/phamclust_pypack.h);
    the plain-Python packer is its executable specification.  Same arrays, dtypes, vocabulary and
    names; errors keep the reference's types (genome.py:39-43); a translation with a character
    outside latin-1 is refused instead of being mangled."""
    import ctypes
    import gc
    from oracle import testdata
    from phamclust_b200 import packing
    from phamclust_b200.genome import Genome
    header = (ROOT / "include" / "phamclust_pypack.h").read_text()
    declared = set(re.findall(r"^\s*(?:int|void \*|void)\s*\**(pcpy_\w+)\s*\(", header, re.M))
    lib = ctypes.PyDLL(str(ROOT / "phamclust_b200" / "libphamclust_pypack.so"))
    assert declared == {"pcpy_pack", "pcpy_buffers", "pcpy_free"} and all(hasattr(lib, n) for n in declared)
    rows = testdata.genome_rows(seed=3, n_genomes=9, n_phams=14, paralog_p=0.4)
    genomes = testdata.rows_to_genomes(rows, cls=Genome)
    genomes.append(Genome("empty"))
    odd = Genome("odd")
    odd.phams["zero_copies"] = []
    odd.add("latin1", "MK\xe9T")
    genomes.append(odd)
    for with_sequences in (True, False):
        a = packing.pack_genomes(genomes, with_sequences)
        b = packing.pack_genomes_py(genomes, with_sequences)
        assert a.names == b.names and a.phams == b.phams
        for key in ("genome_ptr", "pham_id", "copies", "len_sum") + (("seq_off", "residues") if with_sequences else ()):
            x, y = getattr(a, key), getattr(b, key)
            assert x.dtype == y.dtype and np.array_equal(x, y), key
        assert with_sequences or (a.seq_off is None and a.residues is None)
    keep = packing.pack_genomes(genomes).residues          # outlives its table: the buffers stay alive
    gc.collect()
    assert bytes(keep[:3]) == packing.pack_genomes_py(genomes).residues[:3].tobytes()
    assert packing.pack_genomes([]).n_genomes == 0
    assert packing.pack_genomes(iter(genomes[:2])).n_genomes == 2
    wide = Genome("wide")
    wide.add("p", "MK\u0100")
    with pytest.raises(ValueError):
        packing.pack_genomes([wide])
    with pytest.raises(ValueError):
        packing.pack_genomes_py([wide])
    bad = Genome("bad")
    bad.phams[7] = ["M"]
    with pytest.raises(TypeError):
        packing.pack_genomes([bad])
    bad2 = Genome("bad2")
    bad2.phams["p"] = [b"M"]
    with pytest.raises(TypeError):
        packing.pack_genomes([bad2])
    with pytest.raises(AttributeError):
        packing.pack_genomes([object()])


def test_take_genomes_is_a_random_access_subset():
    import random
    from phamclust_b200 import packing, synthetic
    table = synthetic.phage_table(120, seed=5)
    genomes = packing.unpack_genomes(table)
    idx = random.Random(42).sample(range(120), 37)
    sub = packing.take_genomes(table, idx)
    got = packing.unpack_genomes(sub)
    assert [g.name for g in got] == [genomes[i].name for i in idx]
    for k, i in enumerate(idx):
        assert got[k].phams == genomes[i].phams


def test_fasta_dir_streams_into_the_same_table(tmp_path, reference):
    """table_from_fasta_dir (files scanned in the library, no Genome dicts) == the reference
    loader (scripts/phamclust.py:50-83 -> Genome.load -> read_fasta) followed by pack_genomes:
    wrapped and unwrapped sequence lines, CRLF, trailing blanks, extra header fields, paralogs,
    non-FASTA files and sub-directories skipped; the loader's error cases keep their types."""
    import random
    from phamclust_b200 import packing
    from phamclust_b200.genome import GenomeLoadError
    rows = testdata.genome_rows(seed=8, n_genomes=10, n_phams=18, paralog_p=0.4)
    genomes = testdata.rows_to_genomes(rows, cls=Genome)
    rng = random.Random(1)
    for g in genomes:
        with open(tmp_path / f"{g.name}{rng.choice(['.faa', '.fasta', '.fa'])}", "w", newline="") as fh:
            for pham, seqs in g.phams.items():
                for k, sq in enumerate(seqs):
                    nl = rng.choice(["\n", "\r\n"])
                    fh.write(f">name={g.name}|pham={pham}|n={k}  {nl}")
                    width = rng.choice([7, 60, 10 ** 6])
                    for i in range(0, len(sq), width):
                        fh.write(sq[i:i + width] + rng.choice(["", " ", "\t"]) + nl)
    (tmp_path / "notes.txt").write_text("not fasta")
    (tmp_path / "sub.fa").mkdir()
    ours = load_genomes_from_fasta_dir(tmp_path)
    # the reference's Genome.load + read_fasta on the same files (its directory loader sits in
    # scripts/phamclust.py, which needs plotly to import)
    for g in ours:
        theirs = reference.genome.Genome(g.name)
        theirs.load(next(f for f in tmp_path.iterdir() if f.stem == g.name and f.is_file()))
        assert theirs.phams == g.phams
    want = packing.pack_genomes(ours)
    for with_sequences in (True, False):
        got = packing.table_from_fasta_dir(tmp_path, with_sequences)
        assert got.names == want.names and got.phams == want.phams
        for key in ("genome_ptr", "pham_id", "copies", "len_sum") + (("seq_off", "residues") if with_sequences else ()):
            assert np.array_equal(getattr(got, key), getattr(want, key)), key
    bad = tmp_path / "bad"
    bad.mkdir()
    (bad / "a.fa").write_text(">name=a|n=1\nMK\n")
    with pytest.raises(GenomeLoadError):
        packing.table_from_fasta_dir(bad)
    (bad / "a.fa").write_text("MK\n>pham=x\nMK\n")
    with pytest.raises(ValueError):
        packing.table_from_fasta_dir(bad)
    (bad / "a.fa").write_text(">pham=x|junk\nMK\n")          # upstream stops at the first pham= field
    assert packing.table_from_fasta_dir(bad).phams == ["x"]
    (bad / "a.fa").write_text(">junk|pham=x\nMK\n")
    with pytest.raises(ValueError):
        packing.table_from_fasta_dir(bad)
    (bad / "a.fa").write_text(">junk=1=2|pham=x\nMK\n")
    with pytest.raises(ValueError):
        packing.table_from_fasta_dir(bad)
    (bad / "a.fa").write_text(">pham=x\nMK\n")
    (bad / "a.faa").write_text(">pham=x\nMK\n")
    with pytest.raises(ValueError):
        packing.table_from_fasta_dir(bad)


def test_medoid_on_the_packed_triangle_equals_the_reference(reference):
    """medoid / anti_medoid from exact integer row sums of the packed triangle (pc_node_sums) with
    upstream's float arithmetic only among tied nodes: same (name, value) as the verbatim
    reference class (matrix.py:81-127) for distances and similarities, permuted traversal
    order, exact ties (first / last in traversal order as ``sorted`` would) and single nodes."""
    ref_matrix = reference.matrix
    rng = np.random.default_rng(12)
    for n, tie in ((1, False), (2, False), (7, False), (40, False), (25, True), (300, False)):
        names = [f"g{k:03d}" for k in range(n)]
        if tie:
            vals = rng.integers(0, 4, size=(n, n)) * 250000          # lots of equal sums
        else:
            vals = rng.integers(0, 1000001, size=(n, n))
        vals = np.triu(vals, 1)
        vals = vals + vals.T
        for is_distance in (True, False):
            diag = 0 if is_distance else 1000000
            ours = SymMatrix(list(names), is_distance=is_distance)
            theirs = ref_matrix.SymMatrix(list(names), is_distance=is_distance)
            for i in range(n):
                for j in range(i, n):
                    w = (diag if i == j else int(vals[i, j])) / 1e6
                    ours.set_weight(names[i], names[j], w)
                    theirs.set_weight(names[i], names[j], w)
            for order in (list(names), list(rng.permutation(names))):
                if n > 2 and order != names:
                    ours.reorder(list(order))
                    theirs.reorder(list(order))
                assert ours.nodes == theirs.nodes
                assert ours.medoid == theirs.medoid, (n, tie, is_distance)
                assert ours.anti_medoid == theirs.anti_medoid, (n, tie, is_distance)


def test_tree_order_and_large_submatrices_without_the_reference_limits(reference):
    """``reorder()`` with no argument takes the leaf order of a single-linkage dendrogram
    (matrix.py:673-686).  Ours walks the tree with a stack: the order must equal the verbatim
    reference's (scipy's recursive ``dendrogram``) wherever that one works -- ties included -- and
    still come out for a chain-shaped tree deeper than the interpreter's recursion limit, where
    upstream raises RecursionError.  ``extract_submatrix`` of many nodes (row-wise gather, no
    k(k+1)/2 index arrays) must equal the small-block path entry for entry."""
    from phamclust_b200._native import tri_len
    ref_matrix = reference.matrix
    rng = np.random.default_rng(21)
    for n, hi in ((2, 1000001), (3, 5), (17, 1000001), (60, 30), (150, 1000001)):
        names = [f"n{k:03d}" for k in range(n)]
        vals = np.triu(rng.integers(1, hi, size=(n, n)), 1)
        vals = vals + vals.T
        for is_distance in (True, False):
            ours = SymMatrix(list(names), is_distance=is_distance)
            theirs = ref_matrix.SymMatrix(list(names), is_distance=is_distance)
            for i in range(n):
                for j in range(i, n):
                    w = ((0 if is_distance else 1000000) if i == j else int(vals[i, j]) % 1000001) / 1e6
                    ours.set_weight(names[i], names[j], w)
                    theirs.set_weight(names[i], names[j], w)
            ours.reorder()
            theirs.reorder()
            assert ours.nodes == theirs.nodes, (n, hi, is_distance)
    # a chain: points on a line with growing gaps -> single linkage adds one leaf per merge, the tree is
    # 3000 levels deep (scipy's dendrogram, and with it the reference's reorder(), raises RecursionError)
    n = 3000
    pos = np.cumsum(np.arange(1, n + 1)).astype(np.int64)
    i, j = np.triu_indices(n)
    micro = np.minimum(np.abs(pos[i] - pos[j]), 1000000).astype(np.int32)
    chain = SymMatrix.from_micro([f"c{k:04d}" for k in range(n)], micro, True)
    from scipy.cluster.hierarchy import dendrogram, linkage
    with pytest.raises(RecursionError):
        dendrogram(linkage(chain.to_ndarray(condensed=True), method="single"), no_plot=True, get_leaves=True,
                   count_sort="descending")
    chain.reorder()
    assert sorted(chain.nodes) == sorted(f"c{k:04d}" for k in range(n))
    # the bigger child first at every merge: the growing cluster, then the leaf that joined it
    assert chain.nodes[2:] == [f"c{k:04d}" for k in range(2, n)] and set(chain.nodes[:2]) == {"c0000", "c0001"}
    # large block == small-block path
    pick = [f"c{k:04d}" for k in rng.permutation(n)[:700]]
    big = chain.extract_submatrix(pick)
    assert big.nodes == pick and big.micro.shape[0] == tri_len(700)
    for a in range(0, 700, 97):
        small = chain.extract_submatrix(pick[a:a + 40])
        for x in range(len(small.nodes)):
            for y in range(x, len(small.nodes)):
                assert small.get_weight(small.nodes[x], small.nodes[y]) == big.get_weight(small.nodes[x], small.nodes[y])
    assert big.get_weight(pick[3], pick[650]) == chain.get_weight(pick[3], pick[650])
