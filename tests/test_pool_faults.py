"""hfb200_pool failure handling (reference behaviour: /root/reference/data/watchdog.sh:58-83 sets failed inputs
aside and continues, /root/reference/host/src/main.rs:327-330 surfaces the prover's error): injected device faults are retried on a
re-created context, argument errors fail their own job only, stale control groups are re-committed."""
import numpy as np
import pytest
from conftest import SMALL, make_segment


def _jobs(orc, n, po2=12):
    jobs, expect = [], []
    for i in range(n):
        cir, g, code, data = make_segment(orc, SMALL, po2, trace_seed=900 + i)
        jobs.append((po2, g, code, data, 60 + i))
        expect.append(cir.prove(po2, g, code, data, 60 + i)[0])
    return jobs, expect


def _check_device_fault_recovery(pkg, lib, orc, devices):
    jobs, expect = _jobs(orc, 6)
    with pkg.Pool(devices=devices, contexts_per_device=2, max_po2=12, circuit=SMALL, lib=lib, deterministic=True) as pool:
        pool.inject_fault(worker=0, after_jobs=1, kind=0)   # worker 0 proves one job, then "loses" its context on the next
        seals, devs, ms = pool.prove(jobs, 40000)
        assert all(len(a) == len(b) and (a == b).all() for a, b in zip(seals, expect))   # every job still proved, bit-exact
        st = pool.stats()
        assert st["faults"] == 1 and st["retries"] == 1 and st["contexts_recreated"] == 1 and st["contexts_retired"] == 0
        assert sorted(pool.last_attempts) == [1, 1, 1, 1, 1, 2]
        seals2, _, _ = pool.prove(jobs, 40000)               # the re-created context keeps working
        assert all((a == b).all() for a, b in zip(seals2, expect))
        # a fault on every attempt: the job gives up after 3 attempts, the others are proved
        for _ in range(3):
            pass
    with pkg.Pool(devices=devices[:1], contexts_per_device=1, max_po2=12, circuit=SMALL, lib=lib, deterministic=True) as pool:
        pool.inject_fault(0, 0, 0)
        seals, _, _, errs = pool.prove(jobs[:2], 40000, return_errors=True)
        assert errs == [None, None] and pool.last_attempts[0] == 2   # single worker: its own re-created context takes the job again
        assert (seals[0] == expect[0]).all() and (seals[1] == expect[1]).all()


def test_injected_device_fault_is_retried_on_a_fresh_context(pkg, emu_lib, orc):
    _check_device_fault_recovery(pkg, emu_lib, orc, (0, 0))


@pytest.mark.gpu
def test_injected_device_fault_is_retried_on_gpu(pkg, gpu_lib, orc):
    """One context is destroyed and re-created mid-batch on a real device (arena, streams, events and all); seals stay oracle-equal."""
    _check_device_fault_recovery(pkg, gpu_lib, orc, (0, 0))


def test_non_device_failures_fail_their_job_only(pkg, emu_lib, orc):
    jobs, expect = _jobs(orc, 4)
    bad_globals = jobs[1][1].copy(); bad_globals[5] = np.uint32(0xFFFFFFFF)   # INVALID marker: not a field element
    jobs[1] = (jobs[1][0], bad_globals) + jobs[1][2:]
    with pkg.Pool(devices=(0, 0), contexts_per_device=1, max_po2=12, circuit=SMALL, lib=emu_lib, deterministic=True) as pool:
        seals, _, _, errs = pool.prove(jobs, 40000, return_errors=True)
        assert errs[1] is not None and "non-canonical" in errs[1] and [e for i, e in enumerate(errs) if i != 1] == [None] * 3
        assert pool.last_attempts[1] == 1 and pool.stats()["retries"] == 0       # argument errors are never retried
        for i in (0, 2, 3):
            assert (seals[i] == expect[i]).all()
        with pytest.raises(pkg.Hfb200Error, match="non-canonical"):
            pool.prove(jobs, 40000)
        pool.inject_fault(0, 0, 1)                                                # a non-device failure: reported, not retried
        _, _, _, errs = pool.prove(jobs[2:], 40000, return_errors=True)
        assert sum(e is not None for e in errs) == 1 and pool.stats()["retries"] == 0
        with pytest.raises(pkg.Hfb200Error, match="seal buffer too small"):
            pool.prove(jobs[2:], 100)
        with pytest.raises(pkg.Hfb200Error, match="shape"):
            pool.prove([(12, jobs[0][1], jobs[0][2][:, :100], jobs[0][3], 1)], 40000)


def test_reloaded_control_group_is_recommitted(pkg, emu_lib, orc):
    """a control group reloaded for the same po2 (other columns, or the same buffer rewritten) must not keep serving
    the old commitment; witgen / explicit code columns invalidate the resident group too."""
    po2 = 12
    cir, g, code, data = make_segment(orc, SMALL, po2)
    code2 = code.copy(); code2[6, :100] = code[7, :100]       # different control columns: data no longer satisfies them, but the
    ref1 = cir.prove(po2, g, code, data, 3)[0]                # commitment (code_root) must follow the load all the same
    with pkg.Pool(devices=(0,), contexts_per_device=1, max_po2=po2, circuit=SMALL, lib=emu_lib, deterministic=True) as pool:
        pool.load_control(po2, code)
        s1, _, _ = pool.prove([(po2, g, None, data, 3)], 40000)
        assert (s1[0] == ref1).all()
        buf = code.copy()
        pool.load_control(po2, buf)
        buf[...] = code2                                        # caller rewrites the SAME buffer, then reloads it
        pool.load_control(po2, buf)
        s2, _, _ = pool.prove([(po2, g, None, data, 3)], 40000)
        # the seal commits to the NEW control columns: it equals the oracle's seal for (code2, data) and differs from the old one
        assert (s2[0] == cir.prove(po2, g, code2, data, 3)[0]).all() and not (s2[0] == s1[0]).all()
    with pkg.Context(0, po2, SMALL, lib=emu_lib, deterministic=True) as c:
        c.control_root(po2, code)
        c.witgen_synth(po2, 0x48595046, 1)                      # overwrites the resident code columns: the cached group is gone
        with pytest.raises(pkg.Hfb200Error, match="no control group"):
            c.prove_segment(po2, g, None, data, 3)
        c.prove_segment(po2, g, code, data, 3)
        with pytest.raises(pkg.Hfb200Error, match="no control group"):
            c.segment_begin(po2, g, None, data, 3)              # two-phase entry has the same guard
        c.control_root(po2, code)
        assert (c.prove_segment(po2, g, None, data, 3) == ref1).all()


def test_pool_with_data_defined_circuit(pkg, emu_lib, orc):
    """hfb200_pool_create_ir: the multi-GPU pool over a circuit given as data refuses one-shot jobs with the documented reason
    (the accum columns of a data-defined circuit are the caller's: two-phase API), instead of not existing at all."""
    from oracle import synth_ir
    ir = synth_ir.build(SMALL, 0)
    cir, g, code, data = make_segment(orc, SMALL, 12)
    with pkg.Pool(devices=(0,), contexts_per_device=1, max_po2=12, circuit=SMALL, lib=emu_lib, ir=ir, deterministic=True) as pool:
        _, _, _, errs = pool.prove([(12, g, code, data, 1)], 40000, return_errors=True)
        assert errs[0] is not None and "step_accum is the caller's" in errs[0]
