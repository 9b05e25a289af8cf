//! Raw bindings of `include/hfb200.h`.  Error convention is the one of upstream's sys crates
//! (`risc0_sys::ffi_wrap`): NULL = success, otherwise a malloc'd C string released with `hfb200_free_error`.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_uint};

#[repr(C)]
pub struct hfb200_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct hfb200_pool {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct hfb200_circuit_desc {
    pub w_code: u32,
    pub w_data: u32,
    pub w_accum: u32,
    pub flags: u32,
}

/// Circuit as data: upstream's TapSet and PolyExtStepDef flattened to u32 tables (include/hfb200.h).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct hfb200_tap {
    pub group: u32,
    pub offset: u32,
    pub back: u32,
}
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct hfb200_poly_step {
    pub op: u32,
    pub a: u32,
    pub b: u32,
    pub c: u32,
}
#[repr(C)]
pub struct hfb200_circuit_ir {
    pub w_code: u32,
    pub w_data: u32,
    pub w_accum: u32,
    pub n_mix: u32,
    pub taps: *const hfb200_tap,
    pub n_taps: usize,
    pub steps: *const hfb200_poly_step,
    pub n_steps: usize,
    pub ret: u32,
    /// upstream `CircuitImpl::CIRCUIT_INFO` (16 bytes); all zero selects "RV32IM:v2_______"
    pub info: [u8; 16],
}

#[repr(C)]
pub struct hfb200_segment_job {
    pub po2: u32,
    pub globals: *const u32,
    pub code: *const u32,
    pub data: *const u32,
    pub blind_seed: u64,
    pub seal_out: *mut u32,
    pub seal_cap: usize,
    pub seal_words: usize,
    pub error: *const c_char,
    pub device: c_int,
    pub ms: f32,
    pub attempts: c_int,
}

/// What upstream's `ReceiptClaim` boils down to on this path (include/hfb200.h: receipt claims).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct hfb200_claim {
    pub pre: [u32; 8],
    pub post: [u32; 8],
    pub output: [u32; 8],
    pub exit_code: u32,
}
pub const HFB200_EXIT_HALTED: u32 = 0;
pub const HFB200_EXIT_SYSTEM_SPLIT: u32 = 1;
pub const HFB200_BLIND_OS_ENTROPY: c_int = 0;
pub const HFB200_BLIND_DETERMINISTIC: c_int = 1;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct hfb200_pool_stats_t {
    pub contexts: usize,
    pub contexts_retired: usize,
    pub faults: u64,
    pub retries: u64,
    pub contexts_recreated: u64,
}

extern "C" {
    pub fn hfb200_init(device: c_int, max_po2: u32, circuit: *const hfb200_circuit_desc, out: *mut *mut hfb200_ctx) -> *const c_char;
    pub fn hfb200_init_ir(device: c_int, max_po2: u32, circuit: *const hfb200_circuit_ir, out: *mut *mut hfb200_ctx) -> *const c_char;
    /// CUDA source of the eval_check kernel `hfb200_init_ir` specialises with NVRTC (no device needed).
    pub fn hfb200_ir_source(circuit: *const hfb200_circuit_ir, out: *mut c_char, cap: usize, need: *mut usize) -> *const c_char;
    /// 1 when the context's eval_check runs the NVRTC-specialised kernel (0: interpreter kernel / built-in circuit).
    pub fn hfb200_ir_jit_active(ctx: *const hfb200_ctx, compile_ms: *mut f32) -> c_int;
    pub fn hfb200_destroy(ctx: *mut hfb200_ctx);
    pub fn hfb200_free_error(msg: *const c_char);
    pub fn hfb200_seal_words(ctx: *const hfb200_ctx, po2: u32) -> usize;
    pub fn hfb200_prove_segment(
        ctx: *mut hfb200_ctx, po2: u32, globals: *const u32, code: *const u32, data: *const u32, blind_seed: u64,
        seal_out: *mut u32, seal_cap: usize, seal_words: *mut usize,
    ) -> *const c_char;
    pub fn hfb200_segment_begin(
        ctx: *mut hfb200_ctx, po2: u32, globals: *const u32, code: *const u32, data: *const u32, blind_seed: u64,
        mix_out: *mut u32, mix_cap: usize, mix_words: *mut usize,
    ) -> *const c_char;
    pub fn hfb200_segment_finish(ctx: *mut hfb200_ctx, accum_or_null: *const u32, seal_out: *mut u32, seal_cap: usize, seal_words: *mut usize) -> *const c_char;
    pub fn hfb200_pool_create(
        devices: *const c_int, n_devices: c_int, contexts_per_device: c_int, max_po2: u32, circuit: *const hfb200_circuit_desc,
        out: *mut *mut hfb200_pool,
    ) -> *const c_char;
    pub fn hfb200_pool_create_ir(
        devices: *const c_int, n_devices: c_int, contexts_per_device: c_int, max_po2: u32, circuit: *const hfb200_circuit_ir,
        out: *mut *mut hfb200_pool,
    ) -> *const c_char;
    pub fn hfb200_pool_set_blinding(pool: *mut hfb200_pool, mode: c_int) -> *const c_char;
    pub fn hfb200_pool_stats(pool: *const hfb200_pool, out: *mut hfb200_pool_stats_t) -> *const c_char;
    pub fn hfb200_set_blinding(ctx: *mut hfb200_ctx, mode: c_int) -> *const c_char;
    pub fn hfb200_pool_prove(pool: *mut hfb200_pool, jobs: *mut hfb200_segment_job, n_jobs: usize) -> *const c_char;
    pub fn hfb200_digest_bytes(bytes: *const u8, n: usize, out8: *mut u32) -> *const c_char;
    pub fn hfb200_digest_pair(a8: *const u32, b8: *const u32, out8: *mut u32) -> *const c_char;
    pub fn hfb200_claim_encode(claim: *const hfb200_claim, globals: *mut u32) -> *const c_char;
    pub fn hfb200_claim_decode(seal: *const u32, seal_words: usize, out: *mut hfb200_claim) -> *const c_char;
    pub fn hfb200_claim_next_state(pre8: *const u32, index: u32, po2: u32, post8: *mut u32) -> *const c_char;
    pub fn hfb200_verify_claims(
        seals: *const *const u32, seal_words: *const usize, n: usize, image_id8: *const u32, journal: *const u8, journal_len: usize,
    ) -> *const c_char;
    pub fn hfb200_pool_load_control(pool: *mut hfb200_pool, po2: u32, code: *const u32) -> *const c_char;
    pub fn hfb200_pool_destroy(pool: *mut hfb200_pool);
    /// `Receipt::verify` for one segment seal (host code, no device needed).  Exactly one of `circuit` / `ir` is non-null.
    pub fn hfb200_verify_segment(
        circuit: *const hfb200_circuit_desc, ir: *const hfb200_circuit_ir, seal: *const u32, seal_words: usize,
        code_root: *const u32, po2_out: *mut u32,
    ) -> *const c_char;
    /// The same over the `n` segment seals of a composite receipt on up to `threads` host threads (0 = all); reports the first
    /// failing seal.  `code_roots`: n x 8 words, `po2_out`: n words or null, `first_bad`: optional.
    pub fn hfb200_verify_segments(
        circuit: *const hfb200_circuit_desc, ir: *const hfb200_circuit_ir, seals: *const *const u32, seal_words: *const usize, n: usize,
        code_roots: *const u32, po2_out: *mut u32, threads: c_uint, first_bad: *mut usize,
    ) -> *const c_char;
    /// Where the Fiat-Shamir transcript runs: 0 host (default), 1 device, 2 device replayed as a CUDA graph.  Identical seals.
    pub fn hfb200_set_transcript(ctx: *mut hfb200_ctx, mode: c_int) -> *const c_char;
    pub fn hfb200_graph_launches(ctx: *const hfb200_ctx) -> u64;
    /// Control id of (circuit, po2): Merkle root of the committed control columns, computed on the GPU.
    pub fn hfb200_control_root(ctx: *mut hfb200_ctx, po2: u32, code: *const u32, root_out: *mut u32) -> *const c_char;
}

/// Same contract as `risc0_sys::ffi_wrap`.
pub fn ffi_wrap<F: FnOnce() -> *const c_char>(f: F) -> anyhow::Result<()> {
    let e = f();
    if e.is_null() {
        return Ok(());
    }
    let msg = unsafe { std::ffi::CStr::from_ptr(e) }.to_string_lossy().into_owned();
    unsafe { hfb200_free_error(e) };
    Err(anyhow::anyhow!(msg))
}
