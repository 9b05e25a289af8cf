#!/usr/bin/env python3
"""The data-defined circuit path at rv32im-v2 SCALE: a synthetic PolyExtStep list of ~50 k steps over W = 400
columns with ~1100 taps at the library's declared limits (4 back values, 8 tap sets, <= 4 taps per register).

  ir_scale_probe.py compile            (no GPU) source generation + NVRTC compile of the specialised eval_check for sm_100a:
                                       seconds, registers, spill bytes (what hfb200_init_ir pays once per context)
  ir_scale_probe.py run [po2]          (GPU) hfb200_init_ir, one segment through the JIT kernel and one through the interpreter kernel:
                                       init seconds, check-stage ms of both, seals equal (parity with the oracle: tests/test_gpu_parity.py)
Prints one JSON line."""
import ctypes as C
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hfb200_loader
from oracle import synth_ir   # test infrastructure: only builds the tables

pkg = hfb200_loader.load()
N_GROUPS = int(os.environ.get("IR_GROUPS", "560"))


def nvrtc_compile(src):
    lib = C.CDLL("libnvrtc.so.12")
    prog = C.c_void_p()
    assert lib.nvrtcCreateProgram(C.byref(prog), src.encode(), b"eval_check_jit.cu", 0, None, None) == 0
    opts = [b"--gpu-architecture=sm_100a", b"-std=c++17", b"-lineinfo", b"--ptxas-options=-v"]
    arr = (C.c_char_p * len(opts))(*opts)
    t0 = time.time()
    rc = lib.nvrtcCompileProgram(prog, len(opts), arr)
    dt = time.time() - t0
    n = C.c_size_t()
    lib.nvrtcGetProgramLogSize(prog, C.byref(n))
    log = C.create_string_buffer(n.value)
    lib.nvrtcGetProgramLog(prog, log)
    nb = C.c_size_t()
    lib.nvrtcGetCUBINSize(prog, C.byref(nb))
    return rc, dt, log.value.decode(errors="replace"), nb.value


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "compile"
    ir = synth_ir.build_scaled(n_groups=N_GROUPS)
    W = ir["widths"]
    out = {"widths": list(W), "columns": sum(W), "steps": int(len(ir["steps"])), "taps": int(len(ir["taps"])), "constraints": ir["n_constraints"],
           "limits": {"distinct_backs": 4, "distinct_tap_sets": 8, "max_taps_per_register": 4}}
    if mode == "compile":
        t0 = time.time()
        src = pkg.ir_source(ir, W)
        out["source_seconds"] = round(time.time() - t0, 3)
        out["source_bytes"] = len(src)
        rc, dt, log, nb = nvrtc_compile(src)
        out["nvrtc_rc"], out["nvrtc_seconds"], out["cubin_bytes"] = rc, round(dt, 2), nb
        out["ptxas"] = [ln.strip() for ln in log.splitlines() if "registers" in ln or "spill" in ln][:4]
        print(json.dumps(out))
        return
    po2 = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    rng = np.random.default_rng(1)
    n = 1 << po2
    code = rng.integers(0, pkg.P, size=(W[0], n), dtype=np.uint32)
    data = rng.integers(0, pkg.P, size=(W[1], n), dtype=np.uint32)
    accum = rng.integers(0, pkg.P, size=(W[2], n), dtype=np.uint32)
    g = rng.integers(0, pkg.P, size=32, dtype=np.uint32)
    seals = {}
    for jit in ("1", "0"):
        os.environ["HFB200_IR_JIT"] = jit
        t0 = time.time()
        with pkg.Context(0, po2, W, ir=ir, deterministic=True) as c:
            init_s = time.time() - t0
            for _ in range(2):
                c.segment_begin(po2, g, code, data, 1)
                seals[jit] = c.segment_finish(accum)
            st = c.last_stats()
            key = "jit" if jit == "1" else "interpreter"
            out[key] = {"init_seconds": round(init_s, 2), "jit_active": c.ir_jit_active()[0], "compile_ms": round(c.ir_jit_active()[1], 1),
                        "ms_check": round(st["ms_check"], 3), "ms_deep": round(st["ms_deep"], 3), "ms_total": round(st["ms_total"], 2)}
    out["po2"] = po2
    out["seals_equal_jit_vs_interpreter"] = bool((seals["1"] == seals["0"]).all())
    # parity with the CPU oracle at this scale: tests/test_gpu_parity.py::test_data_defined_circuit_at_rv32im_scale_on_gpu
    print(json.dumps(out))


if __name__ == "__main__":
    main()
