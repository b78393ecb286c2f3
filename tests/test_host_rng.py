"""The native MT19937 streams against numpy's legacy RandomState and the golden
KATs recorded from the reference's call sites (CPU only)."""

import numpy as np
import pytest

from tonic_b200.utils.random_state import RandomState


def test_shuffle_kats(golden):
    g = golden('units')
    a = np.arange(8)
    RandomState(0).shuffle(a)
    np.testing.assert_array_equal(a, g['kat3/shuffle8'])
    np.testing.assert_array_equal(a, [6, 2, 1, 7, 3, 0, 5, 4])
    a = np.arange(1000)
    rs = RandomState(123)
    rs.shuffle(a)
    np.testing.assert_array_equal(a, g['kat3/shuffle1000_seed123'])
    rs.shuffle(a)
    np.testing.assert_array_equal(a, g['kat3/shuffle1000_seed123_second'])


def test_randint_kats(golden):
    g = golden('units')
    np.testing.assert_array_equal(RandomState(0).randint(1000, 5), [684, 559, 629, 192, 835])
    rs = RandomState(77)
    np.testing.assert_array_equal(rs.randint(70000, 64), g['kat4/randint_70000_64'])
    np.testing.assert_array_equal(rs.randint(5 * 10 ** 9, 16), g['kat4/randint_5e9_16'])


def test_uniform_normal_kats(golden):
    g = golden('units')
    rs = RandomState(42)
    np.testing.assert_array_equal(rs.uniform(-1, 1, (3, 4)), g['kat6/uniform'])
    np.testing.assert_array_equal(rs.normal((5, 3)), g['kat6/normal'])     # odd count: cached gauss
    np.testing.assert_array_equal(rs.normal((2, 3)), g['kat6/normal_after'])


@pytest.mark.parametrize('seed', [0, 1, 2 ** 32 - 1, 987654321])
def test_against_numpy_live(seed):
    ours, ref = RandomState(seed), np.random.RandomState(seed)
    for n in (1, 2, 17, 4096, 100003):
        a, b = np.arange(n), np.arange(n)
        ours.shuffle(a)
        ref.shuffle(b)
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(ours.randint(n * 7 + 1, 33), ref.randint(n * 7 + 1, size=33))
    np.testing.assert_array_equal(ours.uniform(-1, 1, (5, 3)), ref.uniform(-1, 1, (5, 3)))
    np.testing.assert_array_equal(ours.normal((7, 3)), ref.normal(size=(7, 3)))
    np.testing.assert_array_equal(ours.randint(2 ** 32, 9), ref.randint(2 ** 32, size=9))
    np.testing.assert_array_equal(ours.randint(1, 4), ref.randint(1, size=4))


def test_segment_stream_matches_reference_indices(golden):
    g = golden('units')
    size, workers, iters, bs, seed = g['segidx_c/cfg']
    rs = RandomState(int(seed))
    order = np.arange(size * workers)
    out = []
    for _ in range(iters):
        rs.shuffle(order)
        out.append(order.copy())
    np.testing.assert_array_equal(np.concatenate(out), g['segidx_c/indices'])
