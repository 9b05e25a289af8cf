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

from .binding import Context, Pool, Hfb200Error
from .receipt import Receipt, CompositeReceipt, SegmentReceipt, Journal, encode_journal


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


@dataclass
class ProveInfo:
    receipt: Receipt
    stats: dict = field(default_factory=dict)


class B200Prover:
    """`default_prover()` stand-in: `prove(session) -> ProveInfo` with `.receipt` like the reference call site."""

    def __init__(self, opts: Optional[ProverOpts] = None, lib=None):
        self.opts = opts or ProverOpts()
        self._pool = Pool(devices=self.opts.devices, contexts_per_device=self.opts.contexts_per_device,
                          max_po2=self.opts.max_segment_po2, circuit=self.opts.circuit, lib=lib)
        # seal capacity needs a context-independent formula: ask a throwaway query through the pool's first context
        self._lib = self._pool.lib

    def close(self):
        self._pool.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def prove_segment(self, seg: Segment, seal_cap: int = 1 << 18) -> SegmentReceipt:
        seals, _, _ = self._pool.prove([(seg.po2, seg.globals_, seg.code, seg.data, seg.blind_seed)], seal_cap)
        return SegmentReceipt(seal=seals[0], index=seg.index, hashfn=self.opts.hashfn)

    def prove(self, session: Session, seal_cap: int = 1 << 18) -> ProveInfo:
        for s in session.segments:
            if s.po2 > self.opts.max_segment_po2:
                raise Hfb200Error("segment po2 %d exceeds max_segment_po2 %d" % (s.po2, self.opts.max_segment_po2))
        if self.opts.reuse_control:
            first = {}
            for s in session.segments:
                first.setdefault(s.po2, s.code)
            for po2, code in first.items():
                self._pool.load_control(po2, code)
            jobs = [(s.po2, s.globals_, None, s.data, s.blind_seed) for s in session.segments]
        else:
            jobs = [(s.po2, s.globals_, s.code, s.data, s.blind_seed) for s in session.segments]
        seals, devices, ms = self._pool.prove(jobs, seal_cap)
        segs = [SegmentReceipt(seal=seal, index=s.index, hashfn=self.opts.hashfn) for s, seal in zip(session.segments, seals)]
        receipt = Receipt(CompositeReceipt(segs), Journal(encode_journal(session.journal)))
        return ProveInfo(receipt, {"devices": devices, "segment_ms": ms})


def default_prover(opts: Optional[ProverOpts] = None, lib=None) -> B200Prover:
    return B200Prover(opts, lib=lib)
