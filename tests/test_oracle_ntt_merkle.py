"""Oracle NTT / LDE / Merkle against naive definitions (first-principles known answers)."""
import json
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2013265921


def _rou(k):
    return pow(137, 2**(27 - k), P)


def _brev(i, bits):
    return int(format(i, "0%db" % bits)[::-1], 2) if bits else 0


def test_interpolate_is_inverse_dft(orc):
    rng = np.random.default_rng(0)
    for lg in (1, 3, 6):
        n = 1 << lg
        coeffs = rng.integers(0, P, n).tolist()
        w = _rou(lg)
        evals = [sum(c * pow(w, i * j, P) for j, c in enumerate(coeffs)) % P for i in range(n)]
        got = orc.decode(orc.interpolate_ntt(orc.encode(evals)))
        assert [int(got[_brev(j, lg)]) for j in range(n)] == coeffs  # bit-reversed output order


def test_evaluate_inverts_interpolate_and_bit_reverse(orc):
    rng = np.random.default_rng(1)
    x = rng.integers(0, P, size=(3, 256), dtype=np.uint32)
    assert (orc.evaluate_ntt(orc.interpolate_ntt(x)) == x).all()
    assert (orc.bit_reverse(orc.bit_reverse(x)) == x).all()


def test_lde_is_evaluation_on_shifted_coset(orc):
    rng = np.random.default_rng(2)
    lg, n = 4, 16
    tr = rng.integers(0, P, n).tolist()
    c = orc.zk_shift(orc.interpolate_ntt(orc.encode(tr)))
    lde = orc.decode(orc.expand_ntt(c, 2)).ravel()
    # f = interpolant of the trace on <w_16>; LDE[i] must be f(3 * w_64^i)
    f = orc.decode(orc.bit_reverse(orc.interpolate_ntt(orc.encode(tr)))).tolist()
    w64 = _rou(lg + 2)
    for i in range(4 * n):
        x = 3 * pow(w64, i, P) % P
        assert int(lde[i]) == sum(cj * pow(x, j, P) for j, cj in enumerate(f)) % P


def test_merkle_root_vs_naive(orc):
    rng = np.random.default_rng(3)
    for rows, cols in ((2, 1), (8, 5), (64, 16), (128, 35)):
        m = rng.integers(0, P, size=(cols, rows), dtype=np.uint32)
        level = [orc.hash_elems(m[:, r]) for r in range(rows)]
        root, nodes = orc.merkle(m, True)
        assert all((nodes[rows + r] == level[r]).all() for r in range(rows))
        while len(level) > 1:
            level = [orc.hash_pair(level[2 * i], level[2 * i + 1]) for i in range(len(level) // 2)]
        assert (root == level[0]).all()


def test_golden(orc):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_small.json")))
    x = orc.encode(np.arange(1, 17))
    assert orc.interpolate_ntt(x).tolist() == gold["intt16_of_1_to_16"]
    assert orc.expand_ntt(orc.zk_shift(orc.interpolate_ntt(x)), 2).ravel().tolist() == gold["lde16_of_1_to_16"]
