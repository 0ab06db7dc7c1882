"""``average`` as PhamClust defines it (statistics.py:4-22); used by
``SymMatrix.medoid``.  (On the device the aai mean is the exact integer ratio
sum(matched)/sum(length), see csrc/pc_gather.cu.)"""


def average(values, weights=None):
    if not weights:
        weights = [1] * len(values)
    if len(values) != len(weights):
        raise ValueError(f"got {len(values)} values and {len(weights)} weights")
    return float(sum([x * y for x, y in zip(values, weights)])) / sum(weights)
