"""Circuits as DATA (SURVEY.md section 8b `hfb200_circuit_register`): tap table + constraint polynomial in the shape of
upstream's PolyExtStepDef.  The synthetic circuit is expressed as such data (oracle/synth_ir.py); the oracle interprets
it, the product compiles it to bytecode and runs the interpreter kernel (here: on the host emulator)."""
import numpy as np
import pytest
from conftest import SMALL


def _segment(orc, widths, po2, variant):
    cir = orc.Circuit(*widths, variant=variant)
    code = cir.gen_code(po2)
    g = cir.gen_globals(5)
    data = cir.gen_data(po2, code, g, 5, 1)
    return cir, g, code, data


@pytest.mark.parametrize("variant", [0, 1])
def test_oracle_ir_equals_builtin_formula(orc, variant):
    cir, g, code, data = _segment(orc, SMALL, 12, variant)
    cir_ir = orc.Circuit(*SMALL, variant=variant, use_ir=True)
    seal, cps, _ = cir.prove(12, g, code, data, 1)
    seal_ir, _, _ = cir_ir.prove(12, g, code, data, 1)
    assert (seal == seal_ir).all()
    assert cir_ir.verify(seal, cps["code_root"]) == 12 and cir.verify(seal_ir, cps["code_root"]) == 12
    if variant:
        assert {tuple(t) for t in cir.taps().tolist() if t[2] == 2}  # v2 really taps back 2
        assert cir.n_taps == orc.Circuit(*SMALL).n_taps + SMALL[1] // 4


@pytest.mark.parametrize("variant,nest", [(0, False), (0, True), (1, False), (1, True)])
def test_data_defined_circuit_matches_oracle(pkg, emu_lib, orc, variant, nest):
    from oracle import synth_ir
    widths, po2 = (12, 24, 8), 12
    cir, g, code, data = _segment(orc, widths, po2, variant)
    ir = synth_ir.build(widths, variant, nest=nest)
    cir_ir = orc.Circuit(*widths, variant=variant)
    cir_ir.set_ir(ir["taps"], ir["steps"], ir["ret"], ir.get("info"))
    oseal, ocps, _ = cir_ir.prove(po2, g, code, data, 1)
    with pkg.Context(0, po2, widths, lib=emu_lib, ir=ir) as c:
        mix = c.segment_begin(po2, g, code, data, 1)
        assert (mix == ocps["accum_mix"]).all()
        seal = c.segment_finish(cir.step_accum(po2, data, mix, 1))   # step_accum is the caller's for data-defined circuits
        cps = c.checkpoints()
        for k, v in ocps.items():
            if k in cps:
                assert (cps[k] == v).all(), k
        assert len(seal) == c.seal_words(po2) == len(oseal) and (seal == oseal).all()
        assert cir_ir.verify(seal, ocps["code_root"]) == po2
        if variant == 0 and not nest:
            # the same circuit through the built-in kernels gives the same seal
            with pkg.Context(0, po2, widths, lib=emu_lib) as b:
                assert (b.prove_segment(po2, g, code, data, 1) == seal).all()
        # the library does not own witgen / step_accum of a data-defined circuit
        with pytest.raises(pkg.Hfb200Error, match="caller"):
            c.segment_begin(po2, g, code, data, 1)
            c.segment_finish(None)
        with pytest.raises(pkg.Hfb200Error):
            c.witgen_synth(po2, 1, 1)


def test_ir_validation(pkg, emu_lib):
    from oracle import synth_ir
    ir = synth_ir.build(SMALL, 0)

    def bad(**kw):
        d = dict(ir); d.update(kw)
        with pytest.raises(pkg.Hfb200Error):
            pkg.Context(0, 12, SMALL, lib=emu_lib, ir=d)
    taps = ir["taps"].copy(); taps[[0, 1]] = taps[[1, 0]]
    bad(taps=taps)                                               # not sorted
    taps = ir["taps"].copy(); taps[0, 1] = 999
    bad(taps=taps)                                               # column out of range
    steps = ir["steps"].copy(); steps[5, 0] = 42
    bad(steps=steps)                                             # unknown op
    steps = ir["steps"].copy(); steps[np.argmax(steps[:, 0] == 5), 1] = 10**6
    bad(steps=steps)                                             # operand used before definition
    bad(ret=10**6)                                               # ret is not a mix var
    many = np.array([(2, 0, b) for b in range(6)] + [(2, c, 0) for c in range(1, SMALL[1])], np.uint32)
    bad(taps=many)                                               # > 4 taps on one register


@pytest.mark.parametrize("variant,nest", [(0, False), (1, True)])
def test_jit_source_compiles_for_sm100a(pkg, tmp_path, variant, nest):
    """hfb200_ir_source (no device needed): the eval_check kernel hfb200_init_ir specialises with NVRTC is straight-line
    CUDA that nvcc accepts for sm_100a -- one statement per bytecode instruction, no interpreter loop."""
    import shutil
    import subprocess
    from oracle import synth_ir
    widths = (12, 24, 8)
    ir = synth_ir.build(widths, variant, nest=nest)
    src = pkg.ir_source(ir, widths)
    assert "hfb200_eval_check_jit" in src and "e4a_mac" in src and "switch" not in src
    n_steps = len(ir["steps"])
    body = src[src.index("hfb200_eval_check_jit"):]
    assert body.count(";\n") >= n_steps  # every IR step became a statement
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cu = tmp_path / "jit.cu"
    cu.write_text(src)
    out = tmp_path / "jit.cubin"
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-cubin", "-o", str(out), str(cu)])
    assert out.stat().st_size > 1000
    with pytest.raises(pkg.Hfb200Error):
        bad = dict(ir); bad["ret"] = 10 ** 6
        pkg.ir_source(bad, widths)


def test_chunked_program_equals_the_monolithic_one(pkg, emu_lib, orc, monkeypatch):
    """The constraint polynomial cut into chunks along its top-level AndEqz / AndCond chain (what lets rv32im-sized circuits compile)
    gives the same seal as one program: chunk limit forced down to 64 steps so that the stand-in circuit splits into ~20 chunks,
    including cuts inside runs of AND_COND groups."""
    from oracle import synth_ir
    widths, po2 = (16, 64, 16), 12
    cir = orc.Circuit(*widths, variant=1)
    code = cir.gen_code(po2); g = cir.gen_globals(5); data = cir.gen_data(po2, code, g, 5, 1)
    ir = synth_ir.build(widths, 1, nest=True)
    cir_ir = orc.Circuit(*widths, variant=1)
    cir_ir.set_ir(ir["taps"], ir["steps"], ir["ret"], ir.get("info"))
    oseal, ocps, _ = cir_ir.prove(po2, g, code, data, 1)
    for chunk in ("64", "150", "100000"):
        monkeypatch.setenv("HFB200_IR_CHUNK", chunk)
        with pkg.Context(0, po2, widths, lib=emu_lib, ir=ir) as c:
            mix = c.segment_begin(po2, g, code, data, 1)
            seal = c.segment_finish(cir.step_accum(po2, data, mix, 1))
            assert len(seal) == len(oseal) and (seal == oseal).all(), chunk
        src = pkg.ir_source(ir, widths, lib=emu_lib)
        n_kernels = src.count("__global__")
        assert (n_kernels == 1) == (chunk == "100000") and (chunk != "64" or n_kernels >= 10)


def test_data_defined_circuit_at_rv32im_scale(pkg, emu_lib, orc):
    """the data-defined path at the SCALE of rv32im-v2 -- ~13 k PolyExtSteps here (the 51 k-step variant runs on
    the GPU tier), W = 400 columns, ~1100 taps with 4 distinct back values and 8 distinct tap sets (the library's declared limits):
    the kernels (emulator build) interpreting the chunked bytecode produce the oracle's seal word for word."""
    from oracle import synth_ir
    ir = synth_ir.build_scaled(n_groups=140)
    W, po2 = ir["widths"], 12
    assert len(ir["steps"]) > 12000 and len(ir["taps"]) > 1000
    rng = np.random.default_rng(3)
    cir = orc.Circuit(*W)
    code = cir.gen_code(po2); g = cir.gen_globals(9)
    data = rng.integers(0, orc.P, size=(W[1], 1 << po2), dtype=np.uint32)
    cir_ir = orc.Circuit(*W)
    cir_ir.set_ir(ir["taps"], ir["steps"], ir["ret"], ir.get("info"))
    oseal, ocps, _ = cir_ir.prove(po2, g, code, data, 1)
    with pkg.Context(0, po2, W, lib=emu_lib, ir=ir) as c:
        mix = c.segment_begin(po2, g, code, data, 1)
        assert (mix == ocps["accum_mix"]).all()
        seal = c.segment_finish(cir.step_accum(po2, data, mix, 1))
        assert len(seal) == len(oseal) and (seal == oseal).all()
        assert (c.checkpoint("check_root") == ocps["check_root"]).all()
