"""The int8 tcgen05 engine of the shared-count contraction against the popc
engine and the array oracle: integers bit-exact, every cheap metric."""
import numpy as np
import pytest

from oracle import pham_oracle as po
from phamclust_b200 import synthetic

pytestmark = pytest.mark.gpu

CHEAP = ("gcs", "jc", "pocp", "af")


def _with_lengths(table, seed):
    """Give a sequence-less sparse table varied per-entry lengths (af needs > 1 limb)."""
    rng = np.random.default_rng(seed)
    table.len_sum = (table.copies * rng.integers(30, 1500, size=len(table.copies))).astype(np.uint32)
    return table


@pytest.mark.parametrize("n,p,seed", [(130, 1500, 5), (300, 5000, 6), (257, 40000, 7), (1, 10, 8)])
def test_i8_engine_matches_popc_and_oracle(gpu_ctx, n, p, seed):
    table = _with_lengths(synthetic.sparse_table(n, max(p, 400), seed=seed, clusters=4, mean_phams=30), seed)
    gpu_ctx.load_table(table)
    iu = np.triu_indices(n)
    try:
        for key in CHEAP:
            gpu_ctx.set_engine("popc")
            want_num = gpu_ctx.numerators(key)
            want = {d: gpu_ctx.matrix(key, as_distance=d) for d in (True, False)}
            oracle_num = po.table_numerators(table, key)[iu].copy()
            oracle_num[iu[0] == iu[1]] = 0
            assert np.array_equal(want_num, oracle_num), key
            for engine in ("i8mma", "i8mma_v1"):       # persistent CTA-pair kernel, round-1 kernel
                gpu_ctx.set_engine(engine)
                assert np.array_equal(gpu_ctx.numerators(key), want_num), (key, engine)
                for d in (True, False):
                    assert np.array_equal(gpu_ctx.matrix(key, as_distance=d), want[d]), (key, d, engine)
                # row blocks through the int8 engine recompose as well
                if n > 128:
                    parts = [gpu_ctx.matrix_rows(key, True, r0, r1) for r0, r1 in ((0, 100), (100, 129), (129, n))]
                    assert np.array_equal(np.concatenate(parts), want[True]), (key, engine)
    finally:
        gpu_ctx.set_engine("auto")


def test_i8_engine_at_benchmark_size(gpu_ctx):
    """Many tiles per CTA pair (the persistent kernel's accumulator ring and tile queue wrap many
    times), every metric layout (1 segment, [C|B], two af limbs), against the popc engine; 60 random
    rows of the shared counts also against the array oracle."""
    table = _with_lengths(synthetic.sparse_table(4000, 60000, seed=1), 12)
    gpu_ctx.load_table(table)
    n = table.n_genomes
    try:
        for key in CHEAP:
            gpu_ctx.set_engine("popc")
            want = gpu_ctx.matrix(key)
            for engine in ("i8mma", "i8mma_v1"):
                gpu_ctx.set_engine(engine)
                assert np.array_equal(gpu_ctx.matrix(key), want), (key, engine)
        gpu_ctx.set_engine("i8mma")
        num = gpu_ctx.numerators("gcs")
        dense = po.table_numerators(table, "gcs")
        rng = np.random.default_rng(5)
        for i in rng.choice(n, size=60, replace=False).tolist():
            lo = i * n - (i * (i - 1)) // 2
            got = num[lo:lo + n - i].copy()
            want_row = dense[i, i:].copy()
            want_row[0] = 0
            assert np.array_equal(got, want_row), i
    finally:
        gpu_ctx.set_engine("auto")


@pytest.mark.parametrize("shuffle_ids", [False, True])
def test_i8_block_sparse_schedule_on_clustered_tables(gpu_ctx, shuffle_ids):
    """Genomes of different phage clusters share no K block: most 256 x 256 tiles
    of this table skip the tensor core entirely (shared count 0 by construction),
    the diagonal ones visit a handful of K blocks.  Every value must still equal
    the popc engine's and the oracle's; with shuffled pham ids the library's own
    column order (first genome holding the pham) has to rebuild the blocks."""
    n = 1100                                         # 22 clusters, ragged last tiles
    table = synthetic.phage_table(n, seed=11, with_sequences=False)
    if shuffle_ids:
        perm = np.random.default_rng(3).permutation(table.meta["n_phams"]).astype(np.int32)
        table.pham_id = perm[table.pham_id]
    gpu_ctx.load_table(table)
    iu = np.triu_indices(n)
    try:
        for key in CHEAP:
            gpu_ctx.set_engine("popc")
            want_num = gpu_ctx.numerators(key)
            want = gpu_ctx.matrix(key, as_distance=True)
            oracle_num = po.table_numerators(table, key)[iu].copy()
            oracle_num[iu[0] == iu[1]] = 0
            assert np.array_equal(want_num, oracle_num), key
            gpu_ctx.set_engine("i8mma")
            assert np.array_equal(gpu_ctx.numerators(key), want_num), key
            assert np.array_equal(gpu_ctx.matrix(key, as_distance=True), want), key
            parts = [gpu_ctx.matrix_rows(key, False, r0, r1) for r0, r1 in ((0, 300), (300, 513), (513, n))]
            gpu_ctx.set_engine("popc")
            assert np.array_equal(np.concatenate(parts), gpu_ctx.matrix(key, as_distance=False)), key
    finally:
        gpu_ctx.set_engine("auto")
