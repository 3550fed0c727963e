"""CPU: the numpy restatement of the device RNG (oracle/philox.py) against the published Random123 known-answer
vectors of Philox4x32-10, and its normals against basic distribution properties."""
import numpy as np

from oracle.philox import normal_rows, philox4x32_10

KAT = [  # counter, key, expected (Random123 kat_vectors, philox4x32 10 rounds)
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers():
    for ctr, key, exp in KAT:
        r = philox4x32_10(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))[0]
        assert tuple(int(v) for v in r) == exp


def test_normal_rows_depend_only_on_seed_and_global_row():
    full = normal_rows(7, 0, 1000, 5)
    part = normal_rows(7, 400, 250, 5)
    assert np.array_equal(full[400:650], part)
    assert not np.array_equal(full, normal_rows(8, 0, 1000, 5))
    z = normal_rows(3, 0, 200_000, 4)
    assert np.abs(z.mean(axis=0)).max() < 0.01 and np.abs(z.std(axis=0) - 1).max() < 0.01
    assert np.abs(np.corrcoef(z.T) - np.eye(4)).max() < 0.01
    assert np.abs((z ** 4).mean(axis=0) - 3.0).max() < 0.1
