"""Oracle field arithmetic against first-principles Python integers (SURVEY.md section 8c self-checks)."""
import numpy as np

P = 2013265921
R = 1 << 32
ROU_FWD = [1, 2013265920, 284861408, 1801542727, 567209306, 740045640, 918899846, 1881002012, 1453957774,
           65325759, 1538055801, 515192888, 483885487, 157393079, 1695124103, 2005211659, 1540072241,
           88064245, 1542985445, 1269900459, 1461624142, 825701067, 682402162, 1311873874, 1164520853,
           352275361, 18769, 137]


def test_constants_from_first_principles():
    assert P == 15 * 2**27 + 1
    assert (P * 0x88000001) % R == 1
    assert pow(2, 64, P) == 1172168163
    assert pow(137, 2**27, P) == 1 and pow(137, 2**26, P) != 1
    for k, v in enumerate(ROU_FWD):
        assert pow(137, 2**(27 - k), P) == v
    # -11 is a quadratic non-residue => x^4 + 11 irreducible over Fp (p = 1 mod 4 so -1 is a residue; check x^2+11, x^4+11)
    assert pow(P - 11, (P - 1) // 2, P) == P - 1


def test_mont_mul_add_sub(orc):
    L = orc.lib()
    rng = np.random.default_rng(0)
    rinv = pow(R, P - 2, P)
    for a, b in rng.integers(0, P, size=(2000, 2)).tolist() + [[0, 0], [P - 1, P - 1], [1, P - 1], [0, 5]]:
        assert L.orc_mont_mul(a, b) == a * b * rinv % P
        assert L.orc_add(a, b) == (a + b) % P
        assert L.orc_sub(a, b) == (a - b) % P
    for x in [0, 1, 2, 11, P - 1, 123456789]:
        assert L.orc_encode(x) == x * R % P
        assert L.orc_decode(L.orc_encode(x)) == x
    for k in range(28):
        assert L.orc_decode(L.orc_rou_fwd(k)) == ROU_FWD[k]
        assert L.orc_decode(L.orc_mont_mul(L.orc_rou_fwd(k), L.orc_rou_rev(k))) == 1
    assert (orc.decode(orc.encode(np.arange(1000))) == np.arange(1000)).all()


def _fp4_mul_ref(a, b):
    # schoolbook in Z[x]/(x^4+11)
    r = [0] * 7
    for i in range(4):
        for j in range(4):
            r[i + j] += a[i] * b[j]
    return [(r[k] - 11 * (r[k + 4] if k + 4 < 7 else 0)) % P for k in range(4)]


def test_fp4_mul_inv(orc):
    L = orc.lib()
    rng = np.random.default_rng(1)
    for _ in range(300):
        a = rng.integers(0, P, 4).tolist()
        b = rng.integers(0, P, 4).tolist()
        am, bm = orc.encode(a), orc.encode(b)
        out = np.zeros(4, np.uint32)
        L.orc_fp4_mul(am.ctypes.data, bm.ctypes.data, out.ctypes.data)
        assert orc.decode(out).tolist() == _fp4_mul_ref(a, b)
        inv = np.zeros(4, np.uint32)
        L.orc_fp4_inv(am.ctypes.data, inv.ctypes.data)
        L.orc_fp4_mul(am.ctypes.data, inv.ctypes.data, out.ctypes.data)
        assert orc.decode(out).tolist() == [1, 0, 0, 0]
