"""Product-side segment verifier (`hfb200_verify_segment`, SURVEY.md section 8f row N3: `Receipt::verify` as the
reference calls it at host/src/main.rs:622-624 and verifier/src/main.rs:124-126).  It is host code inside
libhfb200.so and needs no device, so the whole file runs in the CPU tier.  The oracle's verifier is the independent
second implementation: both must accept the same seals and reject the same tampered ones."""
import numpy as np
import pytest
from conftest import SMALL, make_segment


@pytest.fixture(scope="module")
def proved(orc):
    cir, g, code, data = make_segment(orc, SMALL, 12)
    seal, cps, _ = cir.prove(12, g, code, data, 1)
    return cir, seal, cps


def test_accepts_oracle_seal(pkg, gpu_lib, proved):
    cir, seal, cps = proved
    assert pkg.verify_segment(seal, cps["code_root"], SMALL, lib=gpu_lib) == 12
    assert pkg.verify_segment(seal, cir.control_id(12), SMALL, lib=gpu_lib) == 12


@pytest.mark.parametrize("widths,po2", [((16, 64, 16), 12), (SMALL, 13), ((20, 40, 12), 12)])
def test_accepts_other_shapes(pkg, gpu_lib, orc, widths, po2):
    cir, g, code, data = make_segment(orc, widths, po2, trace_seed=9, blind_seed=3)
    seal, cps, _ = cir.prove(po2, g, code, data, 3)
    assert pkg.verify_segment(seal, cps["code_root"], widths, lib=gpu_lib) == po2


def test_tamper_rejected_like_the_oracle(pkg, gpu_lib, proved):
    cir, seal, cps = proved
    rng = np.random.default_rng(1)
    positions = [0, 5, 32, 33, 40, len(seal) // 3, len(seal) // 2, len(seal) - 1] + rng.integers(0, len(seal), 40).tolist()
    for pos in positions:
        bad = seal.copy()
        bad[pos] ^= 1
        with pytest.raises(pkg.Hfb200Error, match="verify"):
            pkg.verify_segment(bad, cps["code_root"], SMALL, lib=gpu_lib)
        with pytest.raises(RuntimeError):
            cir.verify(bad, cps["code_root"])
    with pytest.raises(pkg.Hfb200Error, match="truncated"):
        pkg.verify_segment(seal[:-1], cps["code_root"], SMALL, lib=gpu_lib)
    with pytest.raises(pkg.Hfb200Error, match="trailing"):
        pkg.verify_segment(np.concatenate([seal, seal[:1]]), cps["code_root"], SMALL, lib=gpu_lib)
    wrong = cps["code_root"].copy(); wrong[0] ^= 1
    with pytest.raises(pkg.Hfb200Error, match="control id"):
        pkg.verify_segment(seal, wrong, SMALL, lib=gpu_lib)
    with pytest.raises(pkg.Hfb200Error):
        pkg.verify_segment(seal, cps["code_root"][:7], SMALL, lib=gpu_lib)
    # a seal for another circuit shape cannot be parsed as this one
    with pytest.raises(pkg.Hfb200Error):
        pkg.verify_segment(seal, cps["code_root"], (8, 24, 8), lib=gpu_lib)


def test_random_mutations_never_crash_and_never_verify(pkg, gpu_lib, proved):
    """Robustness: arbitrary word replacements, truncations and extensions end in a clean rejection (the verifier parses
    attacker-controlled data: no out-of-range access, no huge allocation)."""
    cir, seal, cps = proved
    rng = np.random.default_rng(7)
    for trial in range(150):
        bad = seal.copy()
        kind = trial % 4
        if kind == 0:      # one word replaced by a random 32-bit value
            bad[rng.integers(0, len(bad))] = rng.integers(0, 1 << 32, dtype=np.uint64).astype(np.uint32)
        elif kind == 1:    # a run of words replaced
            a = int(rng.integers(0, len(bad) - 64)); bad[a:a + 64] = rng.integers(0, 1 << 31, 64, dtype=np.uint64).astype(np.uint32)
        elif kind == 2:    # truncated anywhere
            bad = bad[:int(rng.integers(0, len(bad)))]
        else:              # the po2 word (seal[32]) set to anything
            bad[32] = np.uint32(rng.integers(0, 64))
            if bad[32] == seal[32]:
                bad[32] = np.uint32(31)
        with pytest.raises(pkg.Hfb200Error):
            pkg.verify_segment(bad, cps["code_root"], SMALL, lib=gpu_lib)
    with pytest.raises(pkg.Hfb200Error):
        pkg.verify_segment(np.zeros(0, np.uint32), cps["code_root"], SMALL, lib=gpu_lib)


def test_non_canonical_element_rejected(pkg, gpu_lib, proved):
    cir, seal, cps = proved
    bad = seal.copy()
    bad[3] = np.uint32(bad[3] + 2013265921) if bad[3] < (1 << 32) - 2013265921 else np.uint32(0xFFFFFFFF)
    with pytest.raises(pkg.Hfb200Error, match="non-canonical"):
        pkg.verify_segment(bad, cps["code_root"], SMALL, lib=gpu_lib)


def test_invalid_witness_rejected(pkg, gpu_lib, orc):
    cir, g, code, data = make_segment(orc, SMALL, 12)
    bad = data.copy()
    bad[cir.w[1] // 2 + 1, 7] ^= 1
    seal, cps, _ = cir.prove(12, g, code, bad, 1)
    with pytest.raises(pkg.Hfb200Error, match="constraint polynomial"):
        pkg.verify_segment(seal, cps["code_root"], SMALL, lib=gpu_lib)


@pytest.mark.parametrize("variant,nest", [(0, False), (1, True)])
def test_data_defined_circuit(pkg, gpu_lib, orc, variant, nest):
    from oracle import synth_ir
    widths, po2 = (12, 24, 8), 12
    cir = orc.Circuit(*widths, variant=variant)
    code = cir.gen_code(po2)
    g = cir.gen_globals(5)
    data = cir.gen_data(po2, code, g, 5, 1)
    ir = synth_ir.build(widths, variant, nest=nest)
    cir.set_ir(ir["taps"], ir["steps"], ir["ret"], ir.get("info"))
    seal, cps, _ = cir.prove(po2, g, code, data, 1)
    assert pkg.verify_segment(seal, cps["code_root"], widths, ir=ir, lib=gpu_lib) == po2
    if variant == 0:
        # same circuit, built-in formula
        assert pkg.verify_segment(seal, cps["code_root"], widths, lib=gpu_lib) == po2
    bad = seal.copy(); bad[len(seal) // 2] ^= 4
    with pytest.raises(pkg.Hfb200Error):
        pkg.verify_segment(bad, cps["code_root"], widths, ir=ir, lib=gpu_lib)
    # a different constraint list must not accept the seal: flip the first Const of the step list
    other = dict(ir)
    steps = np.array(ir["steps"], dtype=np.uint32).reshape(-1, 4).copy()
    k = int(np.nonzero(steps[:, 0] == 0)[0][0])
    steps[k, 1] ^= 1
    other["steps"] = steps.reshape(-1)
    with pytest.raises(pkg.Hfb200Error):
        pkg.verify_segment(seal, cps["code_root"], widths, ir=other, lib=gpu_lib)


def test_emulated_prover_seal_verifies(pkg, emu_lib, gpu_lib, orc):
    """Seal produced by the product's own pipeline (kernel sources on the host emulator) -> product verifier."""
    widths, po2 = SMALL, 12
    cir, g, code, data = make_segment(orc, widths, po2)
    with pkg.Context(0, po2, widths, lib=emu_lib) as c:
        seal = c.prove_segment(po2, g, code, data, 1)
        root = c.control_root(po2, code)
        assert (root == cir.control_id(po2)).all()
    assert pkg.verify_segment(seal, root, widths, lib=gpu_lib) == po2


def test_control_group_cache_on_emulator(pkg, emu_lib, orc):
    """hfb200_control_root keeps the committed control group; code=None reuses it and yields the identical seal."""
    widths, po2 = SMALL, 12
    cir, g, code, data = make_segment(orc, widths, po2)
    _, g2, _, data2 = make_segment(orc, widths, po2, trace_seed=77)
    with pkg.Context(0, 13, widths, lib=emu_lib) as c:
        with pytest.raises(pkg.Hfb200Error, match="no control group"):
            c.prove_segment(po2, g, None, data, 1)
        full = c.prove_segment(po2, g, code, data, 1)
        with pytest.raises(pkg.Hfb200Error, match="no control group"):   # proving does not load the cache by itself
            c.prove_segment(po2, g, None, data, 1)
        c.control_root(po2, code)
        assert (c.prove_segment(po2, g, None, data, 1) == full).all()
        assert (c.prove_segment(po2, g2, None, data2, 5) == cir.prove(po2, g2, code, data2, 5)[0]).all()   # other segment, same control
        mix = c.segment_begin(po2, g, None, data, 1)                        # two-phase form
        with pytest.raises(pkg.Hfb200Error, match="in flight"):
            c.control_root(po2, code)
        assert (c.segment_finish(cir.step_accum(po2, data, mix, 1)) == full).all()
        assert (c.prove_segment(po2, g, code, data, 1) == full).all()      # explicit code drops the cache
        with pytest.raises(pkg.Hfb200Error, match="no control group"):
            c.prove_segment(po2, g, None, data, 1)
        c.control_root(po2, code)
        cir13, g13, code13, data13 = make_segment(orc, widths, 13)
        c.prove_segment(13, g13, code13, data13, 1)                         # another po2 re-lays the arena
        with pytest.raises(pkg.Hfb200Error, match="no control group"):
            c.prove_segment(po2, g, None, data, 1)


@pytest.mark.gpu
def test_gpu_seal_verifies_and_control_root(pkg, orc):
    widths, po2 = (16, 192, 48), 14
    cir, g, code, data = make_segment(orc, widths, po2)
    with pkg.Context(0, po2, widths) as c:
        seal = c.prove_segment(po2, g, code, data, 1)
        root = c.control_root(po2, code)
        assert (root == cir.control_id(po2)).all()
        assert (c.prove_segment(po2, g, None, data, 1) == seal).all()  # cached control group: identical seal
        assert (c.prove_segment(po2, g, code, data, 1) == seal).all()  # control_root leaves the context usable
    assert pkg.verify_segment(seal, root, widths) == po2
    assert cir.verify(seal, root) == po2
    bad = seal.copy(); bad[len(bad) // 2] ^= 1
    with pytest.raises(pkg.Hfb200Error):
        pkg.verify_segment(bad, root, widths)


def test_batch_verifier_fans_out_and_names_the_first_bad_seal(pkg, gpu_lib, orc):
    """hfb200_verify_segments: n seals on several host threads give the same verdicts as n single calls; the FIRST rejected seal
    (lowest index) is the one reported, whatever thread found it."""
    seals, roots = [], []
    for i, po2 in enumerate((12, 13, 12, 12, 13, 12)):
        cir, g, code, data = make_segment(orc, SMALL, po2, trace_seed=100 + i)
        seal, cps, _ = cir.prove(po2, g, code, data, 1)
        seals.append(np.array(seal, np.uint32)); roots.append(np.array(cps["code_root"], np.uint32))
    for threads in (0, 1, 3, 16):
        assert pkg.verify_segments(seals, roots, SMALL, threads=threads, lib=gpu_lib) == [12, 13, 12, 12, 13, 12]
    assert pkg.verify_segments([], [], SMALL, lib=gpu_lib) == []
    bad = [s.copy() for s in seals]
    bad[4][200] ^= 1; bad[2][150] ^= 1
    for threads in (1, 4):
        with pytest.raises(pkg.Hfb200Error, match="segment 2"):
            pkg.verify_segments(bad, roots, SMALL, threads=threads, lib=gpu_lib)
    with pytest.raises(pkg.Hfb200Error, match="segment 1"):                       # a seal checked against another po2's control id
        pkg.verify_segments(seals, [roots[0]] * 6, SMALL, lib=gpu_lib)
    with pytest.raises(pkg.Hfb200Error, match="one 8-word control id per seal"):
        pkg.verify_segments(seals, roots[:3], SMALL, lib=gpu_lib)


def test_transcript_header_binds_proof_system_circuit_and_po2(pkg, gpu_lib, orc):
    """Upstream seeds the transcript with H(PROOF_SYSTEM_INFO), H(CIRCUIT_INFO) and H(globals ++ [po2]) before anything else
    (risc0-circuit-rv32im `SegmentProver::prove`, risc0-zkp `verify`).  The header digest is the hash of the 33-word header; the
    same constraint tables under another CIRCUIT_INFO string are another statement: neither verifier accepts the seal."""
    from oracle import synth_ir
    widths, po2 = (12, 24, 8), 12
    cir = orc.Circuit(*widths)
    code = cir.gen_code(po2); g = cir.gen_globals(5); data = cir.gen_data(po2, code, g, 5, 1)
    seal, cps, _ = cir.prove(po2, g, code, data, 1)
    assert (cps["globals_hash"] == orc.hash_elems(np.concatenate([g, np.array([po2], np.uint32)]))).all()
    ir = synth_ir.build(widths, 0)
    assert ir["info"] == b"SYNTH_RV32IM:v1_"
    assert pkg.verify_segment(seal, cps["code_root"], widths, ir=ir, lib=gpu_lib) == po2       # same circuit as data, same info
    for info in (None, b"RV32IM:v2_______", b"SYNTH_RV32IM:v2_"):                                # None = the data-defined default
        other = dict(ir, info=info)
        with pytest.raises(pkg.Hfb200Error, match="verify"):
            pkg.verify_segment(seal, cps["code_root"], widths, ir=other, lib=gpu_lib)
        c2 = orc.Circuit(*widths)
        c2.set_ir(ir["taps"], ir["steps"], ir["ret"], info)
        with pytest.raises(RuntimeError):
            c2.verify(seal, cps["code_root"])
        s2, cps2, _ = c2.prove(po2, g, code, data, 1)                                            # and its own seals verify under that info only
        assert pkg.verify_segment(s2, cps2["code_root"], widths, ir=other, lib=gpu_lib) == po2
        assert (cps2["code_root"] == cps["code_root"]).all() and not (cps2["accum_mix"] == cps["accum_mix"]).all()
    with pytest.raises(pkg.Hfb200Error, match="16 bytes"):
        pkg.verify_segment(seal, cps["code_root"], widths, ir=dict(ir, info=b"short"), lib=gpu_lib)
