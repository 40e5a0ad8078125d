"""index.verify_buckets: the pickled bucket hash has to be the closed form of genelist.py:367-380,
anything else is refused (the kernels use the closed form)."""
import copy

import numpy as np
import pytest

from te_counter_b200 import index as tindex

BS = tindex.BUCKET_SIZE


def _closed_form(chrom, L, R):
    out = {}
    for i, (c, l, r) in enumerate(zip(chrom, L, R)):
        for b in range((l // BS) * BS, ((r + BS) // BS) * BS, BS):       # genelist.py:372-375
            out.setdefault(c, {}).setdefault(b, []).append(i)
    return out


@pytest.fixture()
def case():
    rng = np.random.default_rng(0)
    n = 400
    chrom = [["1", "2", "X"][int(x)] for x in rng.integers(0, 3, n)]
    L = rng.integers(0, 200000, n)
    R = L + rng.integers(1, 30000, n)
    return chrom, L.tolist(), R.tolist(), _closed_form(chrom, L.tolist(), R.tolist())


def test_closed_form_is_accepted(case):
    chrom, L, R, buckets = case
    tindex.verify_buckets(buckets, chrom, L, R)
    tindex.verify_buckets({}, [], [], [])


def _first(buckets):
    c = next(iter(buckets))
    b = next(iter(buckets[c]))
    return c, b


def test_everything_else_is_refused(case):
    chrom, L, R, buckets = case
    c, b = _first(buckets)
    # an entry missing
    m = copy.deepcopy(buckets)
    m[c][b] = m[c][b][:-1]
    with pytest.raises(ValueError):
        tindex.verify_buckets(m, chrom, L, R)
    # a feature listed twice
    m = copy.deepcopy(buckets)
    m[c][b] = m[c][b] + [m[c][b][0]]
    with pytest.raises(ValueError):
        tindex.verify_buckets(m, chrom, L, R)
    # an id out of range
    m = copy.deepcopy(buckets)
    m[c][b] = m[c][b] + [len(L)]
    with pytest.raises(ValueError):
        tindex.verify_buckets(m, chrom, L, R)
    # a feature in a bucket it does not touch (and one of its own dropped, so the total still matches)
    far = next(i for i in range(len(L)) if chrom[i] == c and not (L[i] // BS * BS <= b <= R[i] // BS * BS))
    m = copy.deepcopy(buckets)
    m[c][b] = m[c][b][:-1] + [far]
    with pytest.raises(ValueError):
        tindex.verify_buckets(m, chrom, L, R)
    # a bucket key that is not a multiple of the bucket size (index written with another bucket_size)
    m = copy.deepcopy(buckets)
    m[c][b + 5000] = m[c].pop(b)
    with pytest.raises(ValueError):
        tindex.verify_buckets(m, chrom, L, R)
    # buckets filed under the wrong chromosome
    m = copy.deepcopy(buckets)
    other = next(k for k in m if k != c)
    m[c], m[other] = m[other], m[c]
    with pytest.raises(ValueError):
        tindex.verify_buckets(m, chrom, L, R)
    # the hash of another bucket size altogether
    with pytest.raises(ValueError):
        tindex.verify_buckets(buckets, chrom, L, R, bs=5000)
