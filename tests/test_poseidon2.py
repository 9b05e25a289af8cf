"""Poseidon2 constants provenance + oracle sponge behaviour."""
import importlib.util
import json
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2013265921
# first row of the Horizen Labs BabyBear t=24 instance as used upstream (recalled; reproduced by the Grain LFSR)
RC_ROW0 = [0x0fa20c37, 0x0795bb97, 0x12c60b9c, 0x0eabd88e, 0x096485ca, 0x07093527, 0x1b1d4e50, 0x30a01ace,
           0x3bd86f5a, 0x69af7c28, 0x3f94775f, 0x731560e8, 0x465a0ecd, 0x574ef807, 0x62fd4870, 0x52ccfe44,
           0x14772b14, 0x4dedf371, 0x260acd7c, 0x1f51dc58, 0x75125532, 0x686a4d7b, 0x54bac179, 0x31947706]


# Upstream's own known-answer test of the permutation: risc0-zkp 3.0.4 `core/hash/poseidon2/mod.rs`, test `poseidon2_test_vectors`
# (/root/reference/Cargo.lock:3195-3198; crate not vendored): poseidon2_mix of the cells [Elem::new(0), ..., Elem::new(23)] read back
# with `as_u32()`.  The 24 words were written down from recollection of that test BEFORE any implementation in this repository was
# run on the input; all 24 matched on the first comparison.  A match on 24 x 31 bits cannot be a coincidence, so it pins everything
# the permutation depends on: the 213 round constants, M_INT_DIAG (until round 2 labelled "recalled, unverified"), the external
# M4 / circulant layer, the 4 + 21 + 4 round structure, the x^7 S-box and the Montgomery encoding convention.
UPSTREAM_KAT_0_TO_23 = [0x2ed3e23d, 0x12921fb0, 0x0e659e79, 0x61d81dc9, 0x32bae33b, 0x62486ae3, 0x1e681b60, 0x24b91325,
                        0x2a2ef5b9, 0x50e8593e, 0x5bc818ec, 0x10691997, 0x35a14520, 0x2ba6a3c5, 0x279d47ec, 0x55014e81,
                        0x5953a67f, 0x2f403111, 0x6b8828ff, 0x1801301f, 0x2749207a, 0x3dc9cf21, 0x3c985ba2, 0x57a99864]


def _gen():
    spec = importlib.util.spec_from_file_location("gen_p2", os.path.join(ROOT, "tools", "gen_poseidon2_consts.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_round_constants_known_answer():
    g = _gen()
    rc = g.grain_constants(1, 0, 31, 24, 8, 21, 8 * 24 + 21, P)
    assert rc[:24] == RC_ROW0
    assert len(rc) == 213 and all(0 <= x < P for x in rc)


def test_inc_files_in_sync_with_generator():
    g = _gen()
    txt = g.render()
    for rel in ("oracle/poseidon2_consts.inc", "hyperfridge-r0_b200/csrc/poseidon2_consts.inc"):
        assert open(os.path.join(ROOT, rel)).read() == txt


def _py_permute(state, g):
    """Independent pure-Python Poseidon2 (canonical integers) to pin the oracle's permutation structure."""
    rc = g.grain_constants(1, 0, 31, 24, 8, 21, 213, P)
    first, part, last = rc[:96], rc[96:117], rc[117:]
    diag = g.M_INT_DIAG
    M4 = [[5, 7, 1, 3], [4, 6, 1, 1], [1, 3, 5, 7], [1, 1, 4, 6]]

    def m_ext(s):
        t = []
        for c in range(0, 24, 4):
            t += [sum(M4[r][k] * s[c + k] for k in range(4)) % P for r in range(4)]
        sums = [sum(t[c + j] for c in range(0, 24, 4)) % P for j in range(4)]
        return [(t[i] + sums[i % 4]) % P for i in range(24)]

    s = m_ext(list(state))
    for r in range(4):
        s = m_ext([pow((s[i] + first[r * 24 + i]) % P, 7, P) for i in range(24)])
    for r in range(21):
        s[0] = pow((s[0] + part[r]) % P, 7, P)
        tot = sum(s) % P
        s = [(tot + diag[i] * s[i]) % P for i in range(24)]
    for r in range(4):
        s = m_ext([pow((s[i] + last[r * 24 + i]) % P, 7, P) for i in range(24)])
    return s


def test_oracle_permutation_matches_pure_python(orc):
    g = _gen()
    rng = np.random.default_rng(3)
    for st in [np.zeros(24, np.int64), np.arange(24), rng.integers(0, P, 24)]:
        got = orc.decode(orc.poseidon2_mix(orc.encode(st)))
        assert got.tolist() == _py_permute(st.tolist(), g)


def test_sponge_edge_cases(orc):
    z = np.zeros(24, np.uint32)
    # empty input == one permutation of the zero state
    assert (orc.hash_elems(np.zeros(0, np.uint32)) == orc.poseidon2_mix(z)[:8]).all()
    x = orc.encode(np.arange(1, 40))
    # exactly one block: no padding permutation
    s = z.copy(); s[:16] = x[:16]
    assert (orc.hash_elems(x[:16]) == orc.poseidon2_mix(s)[:8]).all()
    # 17 elements: second block overwrites cell 0 and ZEROES the rest of the rate, capacity carried
    s2 = orc.poseidon2_mix(s); s2[0] = x[16]; s2[1:16] = 0
    assert (orc.hash_elems(x[:17]) == orc.poseidon2_mix(s2)[:8]).all()
    # hash_pair == permutation of a||b||0
    a, b = x[:8], x[8:16]
    assert (orc.hash_pair(a, b) == orc.poseidon2_mix(s)[:8]).all()


def test_rng(orc):
    d = orc.encode(np.arange(8))
    out = orc.rng_draw(d, 40)
    s = np.zeros(24, np.uint32); s[:8] = d
    s = orc.poseidon2_mix(s)
    assert (out[:16] == s[:16]).all()
    s = orc.poseidon2_mix(s)
    assert (out[16:32] == s[:16]).all()


def test_rng_mix_after_squeeze_permutes_first(orc):
    """Poseidon2Rng::mix (risc0-zkp `poseidon2/rng.rs`): "if switching from squeezing, do a poseidon2 mix" -- a mix that follows
    drawn elements permutes once before the digest is added; two mixes in a row do not."""
    P_ = 2013265921
    d1, d2 = orc.encode(np.arange(8)), orc.encode(np.arange(100, 108))
    def absorb(s, d):
        s = s.copy()
        s[:8] = (s[:8].astype(np.uint64) + d) % P_
        return orc.poseidon2_mix(s.astype(np.uint32))
    s1 = absorb(np.zeros(24, np.uint32), d1)
    assert (orc.rng_mix_draw_mix(d1, 0, d2, 16) == absorb(s1, d2)[:16]).all()                         # no squeeze in between
    for n_skip in (1, 5, 16):
        assert (orc.rng_mix_draw_mix(d1, n_skip, d2, 16) == absorb(orc.poseidon2_mix(s1), d2)[:16]).all()
    assert (orc.rng_mix_draw_mix(d1, 17, d2, 16) == absorb(orc.poseidon2_mix(orc.poseidon2_mix(s1)), d2)[:16]).all()


def test_golden(orc):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_small.json")))
    assert orc.poseidon2_mix(np.zeros(24, np.uint32)).tolist() == gold["poseidon2_zero"]
    assert orc.poseidon2_mix(orc.encode(np.arange(24))).tolist() == gold["poseidon2_iota_mont"]
    assert orc.hash_elems(np.zeros(0, np.uint32)).tolist() == gold["hash_empty"]
    assert orc.hash_elems(orc.encode(np.arange(16))).tolist() == gold["hash_16"]
    assert orc.hash_elems(orc.encode(np.arange(17))).tolist() == gold["hash_17"]


def test_upstream_known_answer_vector(orc, pkg, emu_lib):
    """poseidon2_mix([0..23]) == upstream's `poseidon2_test_vectors` goal: the CPU oracle, the independent pure-Python
    permutation, and the product's kernel source (host emulator build) all reproduce it."""
    enc = np.array(orc.encode(list(range(24))), np.uint32)
    assert [int(v) for v in orc.decode(orc.poseidon2_mix(enc))] == UPSTREAM_KAT_0_TO_23
    assert _py_permute(list(range(24)), _gen()) == UPSTREAM_KAT_0_TO_23
    with pkg.Context(0, 12, (8, 16, 8), lib=emu_lib) as c:
        got = c.op_poseidon2(enc.reshape(1, 24))
    assert [int(v) for v in orc.decode(got[0])] == UPSTREAM_KAT_0_TO_23
