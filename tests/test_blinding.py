"""Zero-knowledge blinding (SURVEY.md Appendix A.1 `Elem::random`): the noise rows come from a 256-bit key expanded by
the ChaCha20 block function.  The C ABI's default key is OS entropy per segment (seals not reproducible, still valid); the
deterministic mode used by every parity test derives the key from the 64-bit blind seed."""
import numpy as np
import pytest
from conftest import SMALL, make_segment


def test_chacha20_block_matches_rfc8439_vector(orc):
    # RFC 8439 section 2.3.2: key 00..1f, nonce 00 00 00 09 00 00 00 4a 00 00 00 00, block counter 1
    key = np.frombuffer(bytes(range(32)), dtype="<u4")
    nonce = np.frombuffer(bytes([0, 0, 0, 9, 0, 0, 0, 0x4a, 0, 0, 0, 0]), dtype="<u4")
    out = orc.chacha20_block(key, 1, nonce)
    expect = [0xe4e7f110, 0x15593bd1, 0x1fdd0f50, 0xc47120a3, 0xc7f4d1c7, 0x0368c033, 0x9aaa2204, 0x4e6cd4c3,
              0x466482d2, 0x09aa9f07, 0x05d7c214, 0xa2028bd9, 0xd19c12b5, 0xb94e16de, 0xe883d0cb, 0x4e3c50a2]
    assert out.tolist() == expect


def test_blind_values_are_field_elements_and_depend_on_every_input(orc):
    base = orc.blind_value(1, 0, 0, 0)
    vals = {base, orc.blind_value(2, 0, 0, 0), orc.blind_value(1, 2, 0, 0), orc.blind_value(1, 0, 1, 0), orc.blind_value(1, 0, 0, 1)}
    assert len(vals) == 5 and all(v < orc.P for v in vals)
    assert orc.blind_value(1, 0, 0, 0) == base


def test_deterministic_mode_matches_oracle_noise_rows(pkg, emu_lib, orc):
    """Product (emulator build of the same sources) == oracle on the blinding rows of DATA (witness stand-in) and ACCUM."""
    po2 = 12
    cir, g, code, data = make_segment(orc, SMALL, po2, blind_seed=77)
    with pkg.Context(0, po2, SMALL, lib=emu_lib, deterministic=True) as c:
        c.witgen_synth(po2, 0x48595046, 77)
        assert (c.read_group(2)[:, -1994:] == data[:, -1994:]).all()
        c.prove_resident(77)
        mix = c.checkpoint("accum_mix")
        assert (c.read_group(0) == cir.step_accum(po2, data, mix, 77)).all()


def test_default_blinding_is_os_entropy(pkg, emu_lib, orc):
    """Without the opt-in, two proofs of the same segment with the same seed differ (fresh key per segment), and both verify."""
    po2 = 12
    cir, g, code, data = make_segment(orc, SMALL, po2)
    with pkg.Context(0, po2, SMALL, lib=emu_lib, deterministic=False) as c:
        a = c.prove_segment(po2, g, code, data, 1)
        b = c.prove_segment(po2, g, code, data, 1)
        assert len(a) == len(b) and not (a == b).all()
        assert (a[:33] == b[:33]).all()                      # same statement
        for seal in (a, b):
            assert cir.verify(seal, cir.control_id(po2)) == po2
        c.set_blinding(pkg.BLIND_DETERMINISTIC)              # opt-in: reproducible, equal to the oracle
        d = c.prove_segment(po2, g, code, data, 1)
        assert (d == c.prove_segment(po2, g, code, data, 1)).all()
        assert (d == cir.prove(po2, g, code, data, 1)[0]).all()
        with pytest.raises(pkg.Hfb200Error):
            c.set_blinding(7)
