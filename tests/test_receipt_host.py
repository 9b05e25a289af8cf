"""The formats either side of the path: journal encoding pinned against the reference's own
receipt fixture, receipt JSON round trip, and the host mirror `default_prover().prove(session).receipt` on the emulator."""
import json
import os
import numpy as np
import pytest
from conftest import SMALL, make_segment

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_journal_encoding_matches_reference_fixture(pkg):
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_journal.json")))
    raw = bytes(fx["journal_bytes"])
    text = pkg.decode_journal(raw)
    commitment = json.loads(text)                       # the guest commits one JSON string
    assert commitment["iban"] == "CH4308307000289537312"  # /root/reference/host/src/main.rs:451-455
    assert [s["elctrnc_seq_nb"] for s in commitment["stmts"]][0] == "247"
    assert pkg.encode_journal(text) == raw              # u32-LE length + utf-8 + zero pad to 4
    r = pkg.Receipt.from_json(json.dumps({"inner": fx["inner"], "journal": {"bytes": fx["journal_bytes"]}}))
    assert r.inner == "Fake" and r.journal.decode() == text
    assert json.loads(r.to_json()) == {"inner": "Fake", "journal": {"bytes": fx["journal_bytes"]}}


def test_host_mirror_proves_a_session(pkg, emu_lib, orc):
    segs, expect = [], []
    for i, po2 in enumerate([12, 12, 13]):
        cir, g, code, data = make_segment(orc, SMALL, po2, trace_seed=500 + i)
        segs.append(pkg.Segment(i, po2, g, code, data, 9 + i))
        expect.append(cir.prove(po2, g, code, data, 9 + i)[0])
    opts = pkg.ProverOpts(max_segment_po2=13, circuit=SMALL, devices=(0, 0), contexts_per_device=1)
    with pkg.default_prover(opts, lib=emu_lib) as prover:
        info = prover.prove(pkg.Session(segs, journal='{"iban":"CH00"}'))
    # opt-in control-group reuse through the pool: same seals (segments of equal po2 share their control columns here)
    opts_rc = pkg.ProverOpts(max_segment_po2=13, circuit=SMALL, devices=(0, 0), contexts_per_device=1, reuse_control=True)
    with pkg.default_prover(opts_rc, lib=emu_lib) as prover:
        info_rc = prover.prove(pkg.Session(segs, journal='{"iban":"CH00"}'))
        with pytest.raises(pkg.Hfb200Error, match="no control group"):
            prover._pool.prove([(14, segs[0].globals_, None, segs[0].data, 1)], 1 << 18)
    for a, b in zip(info.receipt.inner.segments, info_rc.receipt.inner.segments):
        assert np.array_equal(np.asarray(a.seal), np.asarray(b.seal))
    rec = pkg.Receipt.from_json(info.receipt.to_json())      # serde-shaped JSON round trip
    assert [s.index for s in rec.inner.segments] == [0, 1, 2]
    for s, e in zip(rec.inner.segments, expect):
        assert np.array_equal(np.array(s.seal, np.uint32), e)
    assert rec.journal.decode() == '{"iban":"CH00"}'
    cir = orc.Circuit(*SMALL)
    for s, seg in zip(rec.inner.segments, segs):                    # every segment seal verifies
        assert cir.verify(np.array(s.seal, np.uint32), cir.control_id(seg.po2)) == seg.po2
    # `receipt.verify(id)` as the reference's host and verifier call it: product-side verifier, per-po2 control ids
    ids = {12: cir.control_id(12), 13: cir.control_id(13)}
    rec.verify(ids, circuit=SMALL, lib=emu_lib)
    with pytest.raises(pkg.Hfb200Error, match="no control id"):
        rec.verify({12: ids[12]}, circuit=SMALL, lib=emu_lib)
    rec.inner.segments[1].seal[100] ^= 1
    with pytest.raises(pkg.Hfb200Error, match="segment 1"):
        rec.verify(ids, circuit=SMALL, lib=emu_lib)
    rec.inner.segments[1].seal[100] ^= 1
    rec.inner.segments[2].index = 5
    with pytest.raises(pkg.Hfb200Error, match="segment index"):
        rec.verify(ids, circuit=SMALL, lib=emu_lib)
    with pytest.raises(pkg.Hfb200Error, match="Fake"):
        pkg.Receipt("Fake", pkg.Journal(b"")).verify(ids)
    with pytest.raises(pkg.Hfb200Error):
        pkg.ProverOpts(hashfn="sha-256")
