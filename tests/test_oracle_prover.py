"""Oracle prover <-> verifier: soundness round trips, tamper rejection, seal-size model, golden pins."""
import json
import os
import numpy as np
import pytest
from conftest import SMALL, make_segment

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def proved(orc):
    cir, g, code, data = make_segment(orc, SMALL, 12)
    seal, cps, _ = cir.prove(12, g, code, data, 1)
    return cir, g, code, data, seal, cps


def test_verifier_accepts_and_control_id(orc, proved):
    cir, g, code, data, seal, cps = proved
    assert (cir.control_id(12) == cps["code_root"]).all()
    assert cir.verify(seal, cps["code_root"]) == 12


def test_seal_length_model(orc):
    # the model reproduces the reference's published seal-size steps; here it must equal
    # the emitted length exactly.
    for widths, po2 in ((SMALL, 12), (SMALL, 13), ((16, 64, 16), 12)):
        cir, g, code, data = make_segment(orc, widths, po2)
        seal, _, _ = cir.prove(po2, g, code, data, 1)
        assert len(seal) == cir.seal_words_model(po2)


def test_tamper_rejected(orc, proved):
    cir, g, code, data, seal, cps = proved
    rng = np.random.default_rng(0)
    positions = [0, 5, 32, 33, 40, len(seal) // 3, len(seal) // 2, len(seal) - 1] + rng.integers(0, len(seal), 24).tolist()
    for pos in positions:
        bad = seal.copy()
        bad[pos] ^= 1
        with pytest.raises(RuntimeError):
            cir.verify(bad, cps["code_root"])
    with pytest.raises(RuntimeError):
        cir.verify(seal[:-1], cps["code_root"])
    with pytest.raises(RuntimeError):
        cir.verify(np.concatenate([seal, seal[:1]]), cps["code_root"])
    wrong = cps["code_root"].copy(); wrong[0] ^= 1
    with pytest.raises(RuntimeError):
        cir.verify(seal, wrong)


def test_invalid_witness_is_rejected_by_the_verifier(orc):
    # A trace that violates one constraint still yields a seal (the prover cannot tell: the interpolated
    # "check" is simply no longer constraint/vanishing), but the verifier's check at z must fail.
    cir, g, code, data = make_segment(orc, SMALL, 12)
    bad = data.copy()
    bad[cir.w[1] // 2 + 1, 7] ^= 1  # break one derived cell in an active row
    seal, cps, _ = cir.prove(12, g, code, bad, 1)
    with pytest.raises(RuntimeError, match="constraint polynomial"):
        cir.verify(seal, cps["code_root"])


def test_determinism_and_blinding(orc, proved):
    cir, g, code, data, seal, cps = proved
    seal2, _, _ = cir.prove(12, g, code, data, 1)
    assert (seal == seal2).all()
    seal3, cps3, _ = cir.prove(12, g, code, data, 2)  # other accum blinding seed
    assert (cps3["data_root"] == cps["data_root"]).all() and not (cps3["accum_root"] == cps["accum_root"]).all()
    assert cir.verify(seal3, cps["code_root"]) == 12


def test_golden_checkpoints(orc, proved):
    cir, g, code, data, seal, cps = proved
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_small.json")))["segment"]
    assert gold["seal_words"] == len(seal)
    assert orc.hash_elems(seal % orc.P).tolist() == gold["seal_hash"]
    for k, v in gold["checkpoints"].items():
        assert cps[k].tolist() == v, k
