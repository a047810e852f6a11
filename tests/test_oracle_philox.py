"""Known-answer vectors for the Philox4x32-10 restatement (Random123 kat_vectors, philox4x32 10 rounds)."""
import numpy as np

from oracle import philox

KAT = [
    ((0x00000000,) * 4, (0x00000000,) * 2, (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_philox_known_answers():
    for ctr, key, want in KAT:
        got = philox.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert tuple(int(v) for v in got) == want


def test_normals_are_standard_and_shard_invariant():
    z = philox.standard_normals(4096, 16, 0, seed=3, call_index=5)
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01
    a = philox.standard_normals(64, 16, 128, seed=3, call_index=5)
    assert np.array_equal(a, z[128:192])
    b = philox.standard_normals(64, 16, 128, seed=3, call_index=6)
    assert not np.array_equal(a, b)
