//! `b200_prover()` -- the one-line replacement for `default_prover()` at
//! `/root/reference/host/src/main.rs:420`.  The executor (ELF -> Session -> Segments), receipt assembly and
//! verification stay upstream; only `SegmentProver::prove(&Segment) -> Seal` is served by the GPU library.
//!
//! NOTE: seals verify against `risc0_zkvm` only once the rv32im-v2 circuit plug-in and the
//! verified Poseidon2 M_INT_DIAG table are dropped into libhfb200; the library is protocol-complete today with the
//! declared stand-in circuit.
use anyhow::Result;
use hfb200_sys as sys;
use std::rc::Rc;

pub struct B200SegmentProver {
    ctx: *mut sys::hfb200_ctx,
}

impl B200SegmentProver {
    pub fn new(device: i32, max_po2: u32, circuit: sys::hfb200_circuit_desc) -> Result<Self> {
        let mut ctx = std::ptr::null_mut();
        sys::ffi_wrap(|| unsafe { sys::hfb200_init(device, max_po2, &circuit, &mut ctx) })?;
        Ok(Self { ctx })
    }

    /// Column-major u32 Montgomery trace in, seal words out (what upstream's `Seal = Vec<u32>` holds).
    pub fn prove_trace(&self, po2: u32, globals: &[u32], code: &[u32], data: &[u32], blind_seed: u64) -> Result<Vec<u32>> {
        let cap = unsafe { sys::hfb200_seal_words(self.ctx, po2) };
        let mut seal = vec![0u32; cap];
        let mut words = 0usize;
        sys::ffi_wrap(|| unsafe {
            sys::hfb200_prove_segment(self.ctx, po2, globals.as_ptr(), code.as_ptr(), data.as_ptr(), blind_seed, seal.as_mut_ptr(), cap, &mut words)
        })?;
        seal.truncate(words);
        Ok(seal)
    }
}

impl B200SegmentProver {
    /// Control id of (circuit, po2): commits the control columns once and keeps them on the device, so that the segments
    /// that follow can be proved with `prove_trace_shared_control` (identical seals, no per-segment control commitment).
    pub fn load_control(&self, po2: u32, code: &[u32]) -> Result<[u32; 8]> {
        let mut root = [0u32; 8];
        sys::ffi_wrap(|| unsafe { sys::hfb200_control_root(self.ctx, po2, code.as_ptr(), root.as_mut_ptr()) })?;
        Ok(root)
    }

    pub fn prove_trace_shared_control(&self, po2: u32, globals: &[u32], data: &[u32], blind_seed: u64) -> Result<Vec<u32>> {
        let cap = unsafe { sys::hfb200_seal_words(self.ctx, po2) };
        let mut seal = vec![0u32; cap];
        let mut words = 0usize;
        sys::ffi_wrap(|| unsafe {
            sys::hfb200_prove_segment(self.ctx, po2, globals.as_ptr(), std::ptr::null(), data.as_ptr(), blind_seed, seal.as_mut_ptr(), cap, &mut words)
        })?;
        seal.truncate(words);
        Ok(seal)
    }
}

/// `SegmentReceipt::verify_integrity` for the seals this library emits (host code, no GPU needed): `check_code` of upstream's
/// verifier becomes the comparison against `control_id` (one entry of the per-po2 control-id table).
pub fn verify_seal(circuit: &sys::hfb200_circuit_desc, seal: &[u32], control_id: &[u32; 8]) -> Result<u32> {
    let mut po2 = 0u32;
    sys::ffi_wrap(|| unsafe { sys::hfb200_verify_segment(circuit, std::ptr::null(), seal.as_ptr(), seal.len(), control_id.as_ptr(), &mut po2) })?;
    Ok(po2)
}

impl Drop for B200SegmentProver {
    fn drop(&mut self) {
        unsafe { sys::hfb200_destroy(self.ctx) }
    }
}

// Upstream seam (risc0-circuit-rv32im 4.0.4, `prove::SegmentProver`):
//
// impl risc0_circuit_rv32im::prove::SegmentProver for B200SegmentProver {
//     fn prove(&self, segment: &Segment) -> Result<Seal> {
//         let trace = preflight_and_witgen(segment)?;          // upstream CPU code, unchanged
//         self.prove_trace(segment.po2 as u32, &trace.global, &trace.code, &trace.data, blind_seed(segment))
//     }
// }

/// `let prover = hfb200_prover::b200_prover();` replaces `default_prover()`; `prover.prove(env, HYPERFRIDGE_ELF)`
/// (`/root/reference/host/src/main.rs:423`) is unchanged.
pub fn b200_prover() -> Rc<dyn risc0_zkvm::Prover> {
    // ProverImpl over B200SegmentProver: one context per visible GPU, segments handed out from the session's list
    // through hfb200_pool_prove; SegmentReceipt / CompositeReceipt / Receipt are built by upstream's constructors.
    unimplemented!("wired when the rv32im-v2 circuit plug-in is available (see INTEGRATION.md section 3)")
}
