"""Receipt wire format around the hot path -- host-side data classes only.

Mirrors the serde-JSON shape the reference writes and reads:
  * written by `serde_json::to_string(&receipt)` at /root/reference/host/src/main.rs:250-252,
  * journal decoded at /root/reference/host/src/main.rs:258-267 (`receipt.journal.bytes`: u32-LE length, UTF-8 JSON,
    zero padding to 4 bytes),
  * read back by `serde_json::from_slice::<Receipt>` at /root/reference/verifier/src/main.rs:118-119.
The two receipts shipped with the reference are dev-mode fakes (`{"inner":"Fake","journal":{"bytes":[...]}}`,
/root/reference/data/test/test.xml-Receipt-test.json:1); `Receipt.from_json` accepts them, which pins the journal
encoding against reference data.  The Composite form (`{"inner":{"Composite":{"segments":[{"seal":[..u32..],"index":..,
"hashfn":"poseidon2",..}],..}}`) follows risc0-zkvm 3.0.5's `CompositeReceipt` / `SegmentReceipt` field names (crate not
vendored: recollection).  `claim` = {"pre","post","exit_code","output"} as decoded from the seal's
globals (hfb200_claim_decode); `Receipt.verify(image_id, control_ids)` checks the whole chain (see its docstring).
"""
import json
import struct
from dataclasses import dataclass, field
from typing import List, Optional


def encode_journal(text: str) -> bytes:
    """risc0 serde encoding of one committed `String`: u32-LE byte length, the bytes, zero pad to a multiple of 4."""
    raw = text.encode("utf-8")
    pad = (-len(raw)) % 4
    return struct.pack("<I", len(raw)) + raw + b"\x00" * pad


def decode_journal(data: bytes) -> str:
    if len(data) < 4:
        raise ValueError("journal too short")
    (n,) = struct.unpack("<I", data[:4])
    if 4 + n > len(data):
        raise ValueError("journal length prefix exceeds the payload")
    return data[4:4 + n].decode("utf-8")


@dataclass
class Journal:
    bytes_: bytes = b""

    def decode(self) -> str:
        return decode_journal(self.bytes_)

    def to_obj(self):
        return {"bytes": list(self.bytes_)}


@dataclass
class SegmentReceipt:
    seal: List[int]
    index: int
    hashfn: str = "poseidon2"
    verifier_parameters: List[int] = field(default_factory=lambda: [0] * 8)
    claim: Optional[dict] = None

    def to_obj(self):
        return {"seal": [int(x) for x in self.seal], "index": self.index, "hashfn": self.hashfn,
                "verifier_parameters": self.verifier_parameters, "claim": self.claim}


@dataclass
class CompositeReceipt:
    segments: List[SegmentReceipt]
    assumption_receipts: list = field(default_factory=list)
    verifier_parameters: List[int] = field(default_factory=lambda: [0] * 8)

    def to_obj(self):
        return {"segments": [s.to_obj() for s in self.segments], "assumption_receipts": self.assumption_receipts,
                "verifier_parameters": self.verifier_parameters}


@dataclass
class Receipt:
    inner: object  # "Fake" | CompositeReceipt
    journal: Journal
    metadata: Optional[dict] = None

    def to_json(self) -> str:
        inner = self.inner if isinstance(self.inner, str) else {"Composite": self.inner.to_obj()}
        obj = {"inner": inner, "journal": self.journal.to_obj()}
        if self.metadata is not None:
            obj["metadata"] = self.metadata
        return json.dumps(obj, separators=(",", ":"))

    @staticmethod
    def from_json(text: str) -> "Receipt":
        obj = json.loads(text)
        journal = Journal(bytes(obj["journal"]["bytes"]))
        inner = obj["inner"]
        if isinstance(inner, dict) and "Composite" in inner:
            c = inner["Composite"]
            segs = [SegmentReceipt(s["seal"], s["index"], s.get("hashfn", "poseidon2"), s.get("verifier_parameters", [0] * 8), s.get("claim"))
                    for s in c["segments"]]
            inner = CompositeReceipt(segs, c.get("assumption_receipts", []), c.get("verifier_parameters", [0] * 8))
        return Receipt(inner, journal, obj.get("metadata"))

    def seal_bytes(self) -> int:
        return 0 if isinstance(self.inner, str) else sum(4 * len(s.seal) for s in self.inner.segments)

    def verify(self, image_id, control_ids, circuit=(16, 192, 48), ir=None, lib=None) -> None:
        """`receipt.verify(HYPERFRIDGE_ID)` of the reference (/root/reference/host/src/main.rs:622-624,
        /root/reference/verifier/src/main.rs:124-126; the verifier trusts `receipt.journal` after this call):
          1. every segment seal is checked by `hfb200_verify_segment` against the control id of its po2 (`control_ids`:
             {po2: 8 words}, the analogue of upstream's per-po2 control-id table); segment indices run 0..n-1; a dev-mode
             `Fake` receipt is refused exactly like upstream refuses it outside RISC0_DEV_MODE;
          2. the claim chain (`hfb200_verify_claims`): each segment's claim is decoded from its seal's globals, segment 0
             must start from `image_id`, every segment must continue from its predecessor's post-state, only the last one
             halts, and its output digest must equal the digest of `journal.bytes` -- a receipt whose journal was replaced,
             whose segments were swapped or dropped, or that was proved for another image is rejected;
          3. a `claim` object carried in the JSON must equal the one decoded from the seal.
        Raises Hfb200Error with the reason."""
        from .binding import verify_segment, verify_claims, claim_decode, Hfb200Error
        self.verify_seals(control_ids, circuit=circuit, ir=ir, lib=lib)
        segs = self.inner.segments
        verify_claims([s.seal for s in segs], image_id, self.journal.bytes_, lib=lib)
        for i, s in enumerate(segs):
            if s.claim is not None and s.claim != claim_decode(s.seal, lib=lib).to_obj():
                raise Hfb200Error("verify: segment %d: the claim in the receipt differs from the one its seal commits to" % i)

    def verify_seals(self, control_ids, circuit=(16, 192, 48), ir=None, lib=None, threads=0) -> None:
        """Step 1 of verify() alone: seals against control ids and index order.  NOT a substitute for verify(): it says
        nothing about the journal or the image id.  The seals are independent and are checked on `threads` host threads
        (`hfb200_verify_segments`; 0 = all hardware threads)."""
        from .binding import verify_segments, Hfb200Error
        if isinstance(self.inner, str):
            raise Hfb200Error("verify: %s receipt carries no seal (dev-mode receipts are refused)" % self.inner)
        segs = self.inner.segments
        if not segs:
            raise Hfb200Error("verify: composite receipt without segments")
        roots = []
        for want, s in enumerate(segs):
            if s.index != want:
                raise Hfb200Error("verify: segment index %d at position %d" % (s.index, want))
            if s.hashfn != "poseidon2":
                raise Hfb200Error("verify: hash suite %r is not on this path" % s.hashfn)
            if len(s.seal) < 33:
                raise Hfb200Error("verify: segment %d: seal truncated" % want)
            po2 = int(s.seal[32])  # seal layout: 32 globals, po2, ...
            if po2 not in control_ids:
                raise Hfb200Error("verify: segment %d: no control id for po2 %d" % (want, po2))
            roots.append(control_ids[po2])
        verify_segments([s.seal for s in segs], roots, circuit, ir=ir, threads=threads, lib=lib)
