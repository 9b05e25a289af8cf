"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (hyperfridge-r0_b200/) never does.

PARITY UNPINNED for whole seals: the oracle restates risc0-zkp 3.0.4 / risc0-core 3.0.1 from SURVEY.md Appendix A (plus two
corrections from recollection of upstream's sources); the upstream crates are not vendored under
/root/reference, the reference ships no golden seal, and the circuit is a declared stand-in.  Pinned to upstream: the
Poseidon2 permutation (upstream's own known-answer vector `poseidon2_test_vectors`, tests/test_poseidon2.py), the BabyBear
constants (re-derived), and the seal-length model (the reference's five published seal sizes).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_SRCS = ["prover.cpp", "capi.cpp", "fp.h", "poseidon2.h", "poseidon2_consts.inc", "ntt.h", "merkle_iop.h", "blind.h", "circuit.h", "prover.h", "Makefile"]


def build(force=False):
    """Compile oracle/_build/liboracle.so with the Makefile if missing or stale."""
    stale = force or not os.path.exists(_SO)
    if not stale:
        t = os.path.getmtime(_SO)
        stale = any(os.path.getmtime(os.path.join(_HERE, s)) > t for s in _SRCS)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u32p, vp = C.POINTER(C.c_uint32), C.c_void_p
        L.orc_last_error.restype = C.c_char_p
        for name in ("orc_mont_mul", "orc_add", "orc_sub"):
            getattr(L, name).restype = C.c_uint32
            getattr(L, name).argtypes = [C.c_uint32, C.c_uint32]
        for name in ("orc_encode", "orc_decode", "orc_inv", "orc_rou_fwd", "orc_rou_rev"):
            getattr(L, name).restype = C.c_uint32
            getattr(L, name).argtypes = [C.c_uint32]
        L.orc_seal_words_model.restype = C.c_size_t
        L.orc_seal_words_model.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint]
        L.orc_proof_seal_words.restype = C.c_size_t
        L.orc_proof_seal_words.argtypes = [vp]
        L.orc_proof_n_checkpoints.restype = C.c_size_t
        L.orc_proof_n_checkpoints.argtypes = [vp]
        L.orc_proof_checkpoint.restype = C.c_size_t
        L.orc_proof_checkpoint.argtypes = [vp, C.c_size_t, C.c_char_p, C.c_size_t, vp, C.c_size_t]
        L.orc_proof_seal.argtypes = [vp, vp]
        L.orc_proof_times.argtypes = [vp, vp]
        L.orc_proof_free.argtypes = [vp]
        L.orc_rng_bits.restype = C.c_uint32
        L.orc_rng_bits.argtypes = [vp, C.c_uint, C.c_size_t]
        L.orc_rng_draw.argtypes = [vp, C.c_size_t, vp, C.c_size_t]
        L.orc_rng_mix_draw_mix.argtypes = [vp, C.c_size_t, vp, vp, C.c_size_t]
        L.orc_hash_elems.argtypes = [vp, C.c_size_t, vp]
        L.orc_hash_pair.argtypes = [vp, vp, vp]
        L.orc_poseidon2_mix.argtypes = [vp]
        L.orc_fp4_mul.argtypes = [vp, vp, vp]
        L.orc_fp4_inv.argtypes = [vp, vp]
        L.orc_interpolate_ntt.argtypes = [vp, C.c_size_t, C.c_size_t]
        L.orc_evaluate_ntt.argtypes = [vp, C.c_size_t, C.c_size_t, C.c_uint]
        L.orc_zk_shift.argtypes = [vp, C.c_size_t, C.c_size_t]
        L.orc_expand_ntt.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_uint]
        L.orc_bit_reverse.argtypes = [vp, C.c_size_t, C.c_size_t]
        L.orc_merkle.argtypes = [vp, C.c_size_t, C.c_size_t, vp, vp]
        L.orc_circuit_info.argtypes = [C.c_uint32] * 3 + [vp, vp, vp]
        L.orc_gen_code.argtypes = [C.c_uint32] * 3 + [C.c_uint, vp]
        L.orc_gen_globals.argtypes = [C.c_uint32] * 3 + [C.c_uint64, vp]
        L.orc_gen_data.argtypes = [C.c_uint32] * 3 + [C.c_uint, vp, vp, C.c_uint64, C.c_uint64, vp]
        L.orc_step_accum.argtypes = [C.c_uint32] * 3 + [C.c_uint, vp, vp, C.c_uint64, vp]
        L.orc_control_id.argtypes = [C.c_uint32] * 3 + [C.c_uint, vp]
        L.orc_prove_segment.argtypes = [C.c_uint32] * 3 + [C.c_uint, vp, vp, vp, C.c_uint64, C.POINTER(vp)]
        L.orc_verify_segment.argtypes = [C.c_uint32] * 3 + [vp, C.c_size_t, vp, vp]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_circuit_new.restype = vp
        L.orc_circuit_new.argtypes = [C.c_uint32] * 4
        L.orc_circuit_free.argtypes = [vp]
        L.orc_circuit_set_ir.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.c_uint32, C.c_char_p]
        L.orc_h_n_taps.restype = C.c_size_t
        L.orc_h_n_taps.argtypes = [vp]
        L.orc_h_taps.argtypes = [vp, vp]
        L.orc_h_gen_data.argtypes = [vp, C.c_uint, vp, vp, C.c_uint64, C.c_uint64, vp]
        L.orc_h_control_id.argtypes = [vp, C.c_uint, vp]
        L.orc_h_seal_words_model.restype = C.c_size_t
        L.orc_h_seal_words_model.argtypes = [vp, C.c_uint]
        L.orc_h_prove_segment.argtypes = [vp, C.c_uint, vp, vp, vp, C.c_uint64, C.POINTER(vp)]
        L.orc_h_verify_segment.argtypes = [vp, vp, C.c_size_t, vp, vp]
        L.orc_chacha20_block.argtypes = [vp, C.c_uint32, vp, vp]
        L.orc_blind_value.restype = C.c_uint32
        L.orc_blind_value.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        _lib = L
    return _lib


P = 2013265921


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _check(rc):
    if rc != 0:
        raise RuntimeError(lib().orc_last_error().decode())


def encode(x):
    """canonical ints -> Montgomery u32 (vectorised in numpy: x * 2^32 mod p)."""
    x = np.asarray(x, dtype=np.uint64) % P
    return ((x << np.uint64(32)) % np.uint64(P)).astype(np.uint32)


def decode(m):
    m = np.asarray(m, dtype=np.uint64)
    rinv = pow(1 << 32, P - 2, P)
    hi = (m * np.uint64(rinv >> 16)) % np.uint64(P)
    return ((hi * np.uint64(1 << 16) + m * np.uint64(rinv & 0xFFFF)) % np.uint64(P)).astype(np.uint32)


def chacha20_block(key8, counter, nonce3):
    """RFC 8439 section 2.3 block function (the blinding-noise PRF): 16 output words."""
    key8, nonce3 = _u32(key8), _u32(nonce3)
    out = np.zeros(16, np.uint32)
    lib().orc_chacha20_block(_p(key8), counter, _p(nonce3), _p(out))
    return out


def blind_value(seed, group, col, row):
    return lib().orc_blind_value(seed, group, col, row)


def poseidon2_mix(state24):
    s = _u32(state24).copy()
    lib().orc_poseidon2_mix(_p(s))
    return s


def hash_elems(elems):
    e = _u32(elems)
    out = np.zeros(8, np.uint32)
    lib().orc_hash_elems(_p(e), e.size, _p(out))
    return out


def hash_pair(a, b):
    out = np.zeros(8, np.uint32)
    a, b = _u32(a), _u32(b)
    lib().orc_hash_pair(_p(a), _p(b), _p(out))
    return out


def rng_draw(digests, n_out):
    d = _u32(digests).reshape(-1, 8)
    out = np.zeros(n_out, np.uint32)
    lib().orc_rng_draw(_p(d), d.shape[0], _p(out), n_out)
    return out


def rng_mix_draw_mix(d1, n_skip, d2, n_out):
    """Poseidon2Rng: mix(d1), draw n_skip elements, mix(d2), draw n_out elements."""
    d1, d2 = _u32(d1), _u32(d2)
    out = np.zeros(n_out, np.uint32)
    lib().orc_rng_mix_draw_mix(_p(d1), n_skip, _p(d2), _p(out), n_out)
    return out


def interpolate_ntt(cols):
    a = _u32(cols).copy()
    a2 = a.reshape(-1, a.shape[-1])
    lib().orc_interpolate_ntt(_p(a), a2.shape[0], a2.shape[1])
    return a


def evaluate_ntt(cols, expand_bits=0):
    a = _u32(cols).copy()
    a2 = a.reshape(-1, a.shape[-1])
    lib().orc_evaluate_ntt(_p(a), a2.shape[0], a2.shape[1], expand_bits)
    return a


def zk_shift(cols):
    a = _u32(cols).copy()
    a2 = a.reshape(-1, a.shape[-1])
    lib().orc_zk_shift(_p(a), a2.shape[0], a2.shape[1])
    return a


def expand_ntt(cols, bits=2):
    a = _u32(cols)
    a2 = a.reshape(-1, a.shape[-1])
    out = np.zeros((a2.shape[0], a2.shape[1] << bits), np.uint32)
    lib().orc_expand_ntt(_p(out), _p(a), a2.shape[0], a2.shape[1], bits)
    return out


def bit_reverse(cols):
    a = _u32(cols).copy()
    a2 = a.reshape(-1, a.shape[-1])
    lib().orc_bit_reverse(_p(a), a2.shape[0], a2.shape[1])
    return a


def merkle(matrix, want_nodes=False):
    """matrix: [cols, rows] u32 (column-major rows).  Returns root (and heap-layout nodes [2*rows, 8])."""
    m = _u32(matrix)
    cols, rows = m.shape
    root = np.zeros(8, np.uint32)
    nodes = np.zeros((2 * rows, 8), np.uint32) if want_nodes else None
    _check(lib().orc_merkle(_p(m), rows, cols, _p(root), _p(nodes) if want_nodes else None))
    return (root, nodes) if want_nodes else root


class Circuit:
    """The declared synthetic circuit 'synth-rv32im-shape' (oracle/circuit.h): v1 (variant=0) or v2 (variant=1, tap
    sets {0},{0,1},{0,1,2}).  With `use_ir=True` the constraint polynomial is evaluated by interpreting the PolyStep
    list built by oracle/synth_ir.py instead of the built-in formula (same field values, different code path)."""

    def __init__(self, w_code=16, w_data=192, w_accum=48, variant=0, use_ir=False):
        self.w = (w_code, w_data, w_accum)
        self.variant = variant
        self._h = lib().orc_circuit_new(w_code, w_data, w_accum, variant)
        if not self._h:
            raise RuntimeError(lib().orc_last_error().decode())
        nt, nm, nc = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().orc_circuit_info(*self.w, C.byref(nt), C.byref(nm), C.byref(nc)))
        self.n_mix, self.n_constraints = nm.value, nc.value
        self.ir = None
        if use_ir:
            from . import synth_ir
            self.ir = synth_ir.build(self.w, variant)
            self.set_ir(self.ir["taps"], self.ir["steps"], self.ir["ret"], self.ir.get("info"))
        self.n_taps = lib().orc_h_n_taps(self._h)

    def __del__(self):
        try:
            if self._h:
                lib().orc_circuit_free(self._h)
                self._h = None
        except Exception:
            pass

    def set_ir(self, taps, steps, ret, info=None):
        """info: the circuit's 16-byte CIRCUIT_INFO hashed into the transcript header (None = upstream's b"RV32IM:v2_______")."""
        taps, steps = _u32(taps), _u32(steps)
        if info is not None and len(info) != 16:
            raise ValueError("circuit info must be 16 bytes")
        _check(lib().orc_circuit_set_ir(self._h, _p(taps), taps.size // 3, _p(steps), steps.size // 4, ret, info))

    def taps(self):
        out = np.zeros((lib().orc_h_n_taps(self._h), 3), np.uint32)
        lib().orc_h_taps(self._h, _p(out))
        return out

    def gen_code(self, po2):
        out = np.zeros((self.w[0], 1 << po2), np.uint32)
        _check(lib().orc_gen_code(*self.w, po2, _p(out)))
        return out

    def gen_globals(self, seed):
        out = np.zeros(32, np.uint32)
        _check(lib().orc_gen_globals(*self.w, seed, _p(out)))
        return out

    def gen_data(self, po2, code, globals_, trace_seed, blind_seed):
        out = np.zeros((self.w[1], 1 << po2), np.uint32)
        _check(lib().orc_h_gen_data(self._h, po2, _p(code), _p(globals_), trace_seed, blind_seed, _p(out)))
        return out

    def step_accum(self, po2, data, mix, blind_seed):
        out = np.zeros((self.w[2], 1 << po2), np.uint32)
        mix = _u32(mix)
        _check(lib().orc_step_accum(*self.w, po2, _p(data), _p(mix), blind_seed, _p(out)))
        return out

    def control_id(self, po2):
        out = np.zeros(8, np.uint32)
        _check(lib().orc_h_control_id(self._h, po2, _p(out)))
        return out

    def seal_words_model(self, po2):
        return lib().orc_h_seal_words_model(self._h, po2)

    def prove(self, po2, globals_, code, data, blind_seed):
        """Returns (seal u32 array, checkpoints dict name -> u32 array (ordered), stage times dict)."""
        h = C.c_void_p()
        globals_, code, data = _u32(globals_), _u32(code), _u32(data)
        _check(lib().orc_h_prove_segment(self._h, po2, _p(globals_), _p(code), _p(data), blind_seed, C.byref(h)))
        try:
            n = lib().orc_proof_seal_words(h)
            seal = np.zeros(n, np.uint32)
            lib().orc_proof_seal(h, _p(seal))
            cps = {}
            name = C.create_string_buffer(64)
            buf = np.zeros(4096, np.uint32)
            for i in range(lib().orc_proof_n_checkpoints(h)):
                k = lib().orc_proof_checkpoint(h, i, name, 64, _p(buf), buf.size)
                cps[name.value.decode()] = buf[:k].copy()
            t = np.zeros(6, np.float64)
            lib().orc_proof_times(h, _p(t))
            times = dict(zip(("commit", "accum", "check", "deep", "fri", "total"), t.tolist()))
        finally:
            lib().orc_proof_free(h)
        return seal, cps, times

    def verify(self, seal, code_root):
        seal, code_root = _u32(seal), _u32(code_root)
        po2 = C.c_uint()
        _check(lib().orc_h_verify_segment(self._h, _p(seal), seal.size, _p(code_root), C.byref(po2)))
        return po2.value
