"""The library's counter-based generator (csrc/rng.cu) against the published known-answer vectors of Philox4x32-10 (Random123 kat_vectors:
Salmon et al., SC'11), through the C-ABI on the host - no device needed."""
import ctypes as C

from jstsp19_b200 import _lib

KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


def philox_py(ctr, key):
    """Plain restatement of Philox4x32-10 (10 rounds of two 32x32 -> 64 multiplies and key bumps by the Weyl constants)."""
    c, k = list(ctr), list(key)
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return c


def philox_lib(ctr, key):
    c, k, o = (C.c_uint * 4)(*ctr), (C.c_uint * 2)(*key), (C.c_uint * 4)()
    _lib.lib.jstsp_philox4x32_10(c, k, o)
    return list(o)


def test_philox_known_answers():
    for ctr, key, want in KAT:
        assert philox_py(ctr, key) == want
        assert philox_lib(ctr, key) == want


def test_philox_library_matches_restatement_on_trial_counters():
    for t in (0, 1, 887, 10009, 2 ** 33 + 5):
        for stream in range(5):
            for idx in (0, 1, 4095):
                ctr = [idx, stream, t & 0xFFFFFFFF, t >> 32]
                assert philox_lib(ctr, [20190913, 7]) == philox_py(ctr, [20190913, 7])
