"""Further first-principles pins of the oracle (what the new repo must supply itself):
convolution theorem, linearity of the LDE, FRI fold == evaluation-domain fold, barycentric DEEP evaluation."""
import numpy as np
from hypothesis import given, settings, strategies as st

P = 2013265921


def _rou(k):
    return pow(137, 2**(27 - k), P)


def test_convolution_theorem(orc):
    """iNTT(NTT(a) * NTT(b)) is the cyclic convolution of a and b (exercises both transforms and bit reversal)."""
    rng = np.random.default_rng(0)
    n, lg = 64, 6
    a = rng.integers(0, P, n).tolist(); b = rng.integers(0, P, n).tolist()
    conv = [0] * n
    for i in range(n):
        for j in range(n):
            conv[(i + j) % n] = (conv[(i + j) % n] + a[i] * b[j]) % P
    ea = orc.decode(orc.evaluate_ntt(orc.bit_reverse(orc.encode(a)))).astype(object)
    eb = orc.decode(orc.evaluate_ntt(orc.bit_reverse(orc.encode(b)))).astype(object)
    prod = [int(x) * int(y) % P for x, y in zip(ea, eb)]
    got = orc.decode(orc.bit_reverse(orc.interpolate_ntt(orc.encode(prod)))).tolist()
    assert got == conv


@settings(max_examples=20, deadline=None)
@given(st.integers(0, 2**31), st.integers(1, P - 1), st.integers(1, P - 1))
def test_lde_is_linear(seed, alpha, beta):
    import oracle as orc
    rng = np.random.default_rng(seed)
    x = rng.integers(0, P, 128).astype(np.uint64)
    y = rng.integers(0, P, 128).astype(np.uint64)
    comb = (alpha * x.astype(object) + beta * y.astype(object)) % P

    def lde(v):
        return orc.decode(orc.expand_ntt(orc.zk_shift(orc.interpolate_ntt(orc.encode(np.array(v, dtype=np.uint64)))), 2)).ravel().astype(object)
    assert ((alpha * lde(x) + beta * lde(y)) % P == lde(np.array(comb, dtype=np.uint64))).all()


def test_field_axioms_on_extension(orc):
    L = orc.lib()
    rng = np.random.default_rng(5)

    def mul(a, b):
        out = np.zeros(4, np.uint32)
        a = np.ascontiguousarray(a, np.uint32); b = np.ascontiguousarray(b, np.uint32)
        L.orc_fp4_mul(a.ctypes.data, b.ctypes.data, out.ctypes.data)
        return out
    for _ in range(50):
        a, b, c = (orc.encode(rng.integers(0, P, 4)) for _ in range(3))
        assert (mul(mul(a, b), c) == mul(a, mul(b, c))).all()
        assert (mul(a, b) == mul(b, a)).all()
        s = ((b.astype(np.uint64) + c) % P).astype(np.uint32)
        assert (mul(a, s) == ((mul(a, b).astype(np.uint64) + mul(a, c)) % P).astype(np.uint32)).all()
    x = orc.encode([0, 1, 0, 0])  # x^4 = -11
    x4 = mul(mul(x, x), mul(x, x))
    assert orc.decode(x4).tolist() == [P - 11, 0, 0, 0]


def test_seal_structure_offsets(orc):
    """Walks a seal with the layout of SURVEY.md Appendix B and checks the positions the transcript fixes."""
    from conftest import SMALL, make_segment
    cir, g, code, data = make_segment(orc, SMALL, 13)
    seal, cps, _ = cir.prove(13, g, code, data, 1)
    assert (seal[:32] == g).all() and seal[32] == 13
    pos = 33
    for name in ("code_root", "data_root"):
        tops = seal[pos:pos + 32 * 8].reshape(32, 8)
        level = [t for t in tops]
        while len(level) > 1:
            level = [orc.hash_pair(level[2 * i], level[2 * i + 1]) for i in range(len(level) // 2)]
        assert (level[0] == cps[name]).all()
        pos += 32 * 8
    # the final FRI coefficients hash to the committed digest: they sit right before the 50 query openings
    W = sum(SMALL)
    n_final = (1 << 13) // 16 // 16  # two rounds: 8192 -> 512 -> 32 ... po2 = 13 has 2 rounds
    assert "fri_root_1" in cps and "fri_root_2" not in cps
    per_query = W + 16 + 4 * 8 * (15 - 5) + (64 + 8 * (11 - 5)) + (64 + 8 * (7 - 5))
    fin = seal[len(seal) - 50 * per_query - 4 * n_final:len(seal) - 50 * per_query]
    assert (orc.hash_elems(fin) == cps["fri_final_hash"]).all()
