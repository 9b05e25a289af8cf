"""Host-side mirror of the reference's prover surface for this path (Python, because no Rust toolchain is available;
the Rust binding a maintainer adds is in rust/ and INTEGRATION.md).

Reference surface (/root/reference/host/src/main.rs:389-423):
    let env = ExecutorEnv::builder().write(..)...build();
    let prover = default_prover();
    let receipt = prover.prove(env, HYPERFRIDGE_ELF)?.receipt;
Here the executor (ELF -> Session -> Segments) is out of scope, so the `env` is a
`Session`: the list of segment traces the executor + witness generator would have produced, plus the journal the
guest committed.  Everything from there on -- per-segment proving on the GPU(s), receipt assembly -- has the
reference's names and argument meaning.
"""
from dataclasses import dataclass, field
from typing import List, Optional
import numpy as np

from .binding import (Context, Pool, Hfb200Error, digest_bytes, claim_next_state, claim_encode, claim_decode,
                      EXIT_HALTED, EXIT_SYSTEM_SPLIT)
from .receipt import Receipt, CompositeReceipt, SegmentReceipt, Journal, encode_journal

# Stand-in for HYPERFRIDGE_ID (/root/reference/host/src/main.rs:7, the image id of the guest ELF): the executor is out of scope,
# so there is no ELF to digest; sessions carry this constant unless the caller supplies another id.
IMAGE_ID_TAG = b"hyperfridge guest image id (stand-in: executor out of scope)"


def default_image_id(lib=None):
    return digest_bytes(IMAGE_ID_TAG, lib=lib)


@dataclass
class ProverOpts:
    """Subset of risc0_zkvm::ProverOpts that matters on this path (the host passes none: defaults only)."""
    hashfn: str = "poseidon2"
    receipt_kind: str = "composite"
    max_segment_po2: int = 20
    circuit: tuple = (16, 192, 48)
    devices: tuple = (0,)
    contexts_per_device: int = 2
    reuse_control: bool = False  # opt-in: segments of equal po2 share the control group of the first one (true for rv32im)
    deterministic_blinding: bool = False  # tests / bench only: blinding from Segment.blind_seed alone (NOT zero-knowledge)

    def __post_init__(self):
        if self.hashfn != "poseidon2":
            raise Hfb200Error("only the default poseidon2 hash suite is on the hot path (sha-256 suite: out of scope)")
        if self.receipt_kind != "composite":
            raise Hfb200Error("succinct / groth16 receipts need the recursion circuit: out of scope")


@dataclass
class Segment:
    """What upstream's `Segment` boils down to at the prover seam: po2 and the witness columns (+ blinding seed)."""
    index: int
    po2: int
    globals_: np.ndarray
    code: np.ndarray
    data: np.ndarray
    blind_seed: int


@dataclass
class Session:
    segments: List[Segment]
    journal: str = ""
    image_id: Optional[np.ndarray] = None  # 8 words; None = default_image_id()


@dataclass
class ProveInfo:
    receipt: Receipt
    stats: dict = field(default_factory=dict)


def bind_claims(session: "Session", n_total: Optional[int] = None, lib=None):
    """What upstream's executor does for the prover: every segment's globals get its claim (pre/post state chained from
    the image id, SystemSplit for all but the last segment, the journal digest as the last one's output).  Returns the
    list of per-segment globals (copies; the session is not modified) and the final post-state.
    `n_total`: segments of the whole session when `session` holds only a share of it (multi-rank operation): the chain is a
    hash chain over segment indices from the image id, so every rank computes the same states without communication."""
    image_id = session.image_id if session.image_id is not None else default_image_id(lib)
    journal_digest = digest_bytes(encode_journal(session.journal), lib=lib)
    if n_total is None:
        n_total = max(s.index for s in session.segments) + 1 if session.segments else 0
    po2_of = {s.index: s.po2 for s in session.segments}
    default_po2 = session.segments[0].po2 if session.segments else 0
    states = [np.asarray(image_id, np.uint32)]
    for i in range(n_total):
        states.append(claim_next_state(states[-1], i, po2_of.get(i, default_po2), lib=lib))
    out = []
    for s in session.segments:
        if s.index >= n_total:
            raise Hfb200Error("segment index %d outside the session" % s.index)
        last = s.index == n_total - 1
        out.append(claim_encode(s.globals_, states[s.index], states[s.index + 1], EXIT_HALTED if last else EXIT_SYSTEM_SPLIT,
                                journal_digest if last else np.zeros(8, np.uint32), lib=lib))
    return out, states[-1]


class B200Prover:
    """`default_prover()` stand-in: `prove(session) -> ProveInfo` with `.receipt` like the reference call site."""

    def __init__(self, opts: Optional[ProverOpts] = None, lib=None):
        self.opts = opts or ProverOpts()
        self._pool = Pool(devices=self.opts.devices, contexts_per_device=self.opts.contexts_per_device,
                          max_po2=self.opts.max_segment_po2, circuit=self.opts.circuit, lib=lib,
                          deterministic=self.opts.deterministic_blinding)
        # seal capacity needs a context-independent formula: ask a throwaway query through the pool's first context
        self._lib = self._pool.lib

    def close(self):
        self._pool.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def prove_segment(self, seg: Segment, seal_cap: int = 1 << 18) -> SegmentReceipt:
        """One segment with the globals as given (no claim binding: the caller is the executor)."""
        seals, _, _ = self._pool.prove([(seg.po2, seg.globals_, seg.code, seg.data, seg.blind_seed)], seal_cap)
        return SegmentReceipt(seal=seals[0], index=seg.index, hashfn=self.opts.hashfn)

    def bind_claims(self, session: Session, n_total: Optional[int] = None):
        return bind_claims(session, n_total=n_total, lib=self._lib)

    def prove(self, session: Session, seal_cap: int = 1 << 18, n_total: Optional[int] = None) -> ProveInfo:
        """`prover.prove(env, elf)`: proves every segment of the session (independent jobs over the pool), attaches each
        segment's claim and returns the receipt.  `n_total`: segments of the whole session when this call proves only a share
        of it (multi-rank operation); default = the session is complete."""
        for s in session.segments:
            if s.po2 > self.opts.max_segment_po2:
                raise Hfb200Error("segment po2 %d exceeds max_segment_po2 %d" % (s.po2, self.opts.max_segment_po2))
        globals_with_claims, _ = self.bind_claims(session, n_total=n_total)
        if self.opts.reuse_control:
            first = {}
            for s in session.segments:
                if s.po2 in first:
                    if first[s.po2] is not s.code and not np.array_equal(first[s.po2], s.code):
                        raise Hfb200Error("reuse_control: segments of po2 %d carry different control columns" % s.po2)
                else:
                    first[s.po2] = s.code
            for po2, code in first.items():
                self._pool.load_control(po2, code)
            jobs = [(s.po2, g, None, s.data, s.blind_seed) for s, g in zip(session.segments, globals_with_claims)]
        else:
            jobs = [(s.po2, g, s.code, s.data, s.blind_seed) for s, g in zip(session.segments, globals_with_claims)]
        seals, devices, ms = self._pool.prove(jobs, seal_cap)
        segs = [SegmentReceipt(seal=seal, index=s.index, hashfn=self.opts.hashfn, claim=claim_decode(seal, lib=self._lib).to_obj())
                for s, seal in zip(session.segments, seals)]
        receipt = Receipt(CompositeReceipt(segs), Journal(encode_journal(session.journal)))
        return ProveInfo(receipt, {"devices": devices, "segment_ms": ms, "attempts": list(self._pool.last_attempts)})


def default_prover(opts: Optional[ProverOpts] = None, lib=None) -> B200Prover:
    return B200Prover(opts, lib=lib)
