"""CPU tests for bench.py helpers."""
import numpy as np


def test_synth_generator_matches_oracle(oracle):
    import bench

    n = 100_003
    ref = oracle.synth_packed(bench.SEED, n)
    got, off = bench.synth_packed_range(bench.SEED, 0, n)
    assert off == 0
    nb = (n + 3) // 4
    # the oracle zeroes the unused bits of the last byte; compare whole bytes before it
    assert np.array_equal(got[:nb - 1], ref[:nb - 1])
    # a range starting mid-word addresses the same bases
    got2, off2 = bench.synth_packed_range(bench.SEED, 777, 5000)
    for i in range(0, 4000, 37):
        p1, p2 = 777 + i, off2 + i
        b1 = (int(ref[p1 >> 2]) >> (2 * (p1 & 3))) & 3
        b2 = (int(got2[p2 >> 2]) >> (2 * (p2 & 3))) & 3
        assert b1 == b2
