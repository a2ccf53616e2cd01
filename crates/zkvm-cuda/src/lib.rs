//! Safe Rust wrapper over the C ABI of the B200 Ajtai commitment engine (`include/lattice_ajtai.h`).
//!
//! `latticefold` is `#![forbid(unsafe_code)]` (crates/latticefold/src/lib.rs:4), so all `unsafe` lives here and
//! `latticefold` / `zkvm` call the safe API below.  A `GoldilocksRingNTT` / `GoldilocksRingPoly` is exactly 24
//! contiguous `u64` Montgomery limbs (stark-rings ring/src/cyclotomic_ring/flatten.rs:10-17,
//! models/goldilocks/utils.rs:5-23), so slices are passed by pointer with `LAT_REPR_MONTGOMERY` and no conversion.
use cyclotomic_rings::rings::{GoldilocksRingNTT as NTT, GoldilocksRingPoly as Coeff};
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct LatAjtai {
    _private: [u8; 0],
}

pub const LAT_REPR_MONTGOMERY: c_int = 1;

extern "C" {
    fn lat_strerror(status: c_int) -> *const c_char;
    fn lat_last_error() -> *const c_char;
    fn lat_ajtai_create(out: *mut *mut LatAjtai, kappa: u32, n: u64, log2_b: u32, l: u32, k: u32, repr: c_int, device: c_int) -> c_int;
    fn lat_ajtai_destroy(h: *mut LatAjtai);
    fn lat_ajtai_upload_rows(h: *mut LatAjtai, row0: u32, nrows: u32, rows: *const u64, row_stride: u64) -> c_int;
    fn lat_ajtai_commit_ntt(h: *mut LatAjtai, f: *const u64, f_len: u64, cm: *mut u64) -> c_int;
    fn lat_ajtai_commit_ntt_batch(h: *mut LatAjtai, fs: *const u64, count: u32, f_len: u64, cms: *mut u64) -> c_int;
    fn lat_ajtai_witness_from_w_ccs(h: *mut LatAjtai, w_ccs: *const u64, w_len: u64, f_coeff: *mut u64, f: *mut u64, cm: *mut u64) -> c_int;
    fn lat_ajtai_decompose_commit(h: *mut LatAjtai, f_coeff: *const u64, n: u64, cm: *const u64, planes_coeff: *mut u64, planes_f: *mut u64, cms: *mut u64) -> c_int;
    fn lat_ajtai_submit_w_ccs(h: *mut LatAjtai, w_ccs: *const u64, w_len: u64, cm: *mut u64, ticket: *mut u64) -> c_int;
    fn lat_ajtai_wait(h: *mut LatAjtai, ticket: u64) -> c_int;
    fn lat_ajtai_select_side(h: *mut LatAjtai, side: c_int) -> c_int;
    fn lat_ajtai_fold_witness(h: *mut LatAjtai, rho: *const u64, f0: *mut u64, f0_coeff: *mut u64) -> c_int;
    fn lat_ring_gadget_recompose(f: *const u64, count: u64, log2_b: u32, l: u32, out: *mut u64, repr: c_int, device: c_int) -> c_int;
    fn lat_ring_crt(coeff: *const u64, count: u64, ntt: *mut u64, device: c_int) -> c_int;
    fn lat_ring_icrt(ntt: *const u64, count: u64, coeff: *mut u64, device: c_int) -> c_int;
}

/// Mirrors latticefold::commitment::CommitmentError (crates/latticefold/src/commitment.rs:13-26) plus engine errors.
#[derive(Debug, thiserror::Error)]
pub enum CudaCommitError {
    #[error("Wrong length of the witness: {0}, expected: {1}")]
    WrongWitnessLength(usize, usize),
    #[error("a coefficient needs more digits than the decomposition padding (the CPU path panics here)")]
    DigitOverflow,
    #[error("engine failure {0}: {1}")]
    Engine(i32, String),
}

fn check(status: c_int, got: usize, expected: usize) -> Result<(), CudaCommitError> {
    match status {
        0 => Ok(()),
        1 => Err(CudaCommitError::WrongWitnessLength(got, expected)),
        4 => Err(CudaCommitError::DigitOverflow),
        s => {
            let msg = unsafe { std::ffi::CStr::from_ptr(lat_last_error()) }.to_string_lossy().into_owned();
            let _ = unsafe { lat_strerror(s) };
            Err(CudaCommitError::Engine(s, msg))
        }
    }
}

fn limbs<T>(v: &[T]) -> *const u64 {
    v.as_ptr() as *const u64 // T is 24 contiguous Montgomery u64 (see module docs)
}
fn limbs_mut<T>(v: &mut [T]) -> *mut u64 {
    v.as_mut_ptr() as *mut u64
}

/// Device-resident replacement of `AjtaiCommitmentScheme<GoldilocksRingNTT>`.
pub struct CudaAjtai {
    h: *mut LatAjtai,
    kappa: usize,
    n: usize,
    k: usize,
}
// one caller at a time (ZKVM/main.rs:121-219 is single-threaded); the handle itself is just a pointer
unsafe impl Send for CudaAjtai {}
unsafe impl Sync for CudaAjtai {}

impl CudaAjtai {
    /// AjtaiCommitmentScheme::new(Matrix<R>) (commitment_scheme.rs:49): uploads the host's rows as they are
    /// (`Vec<Vec<R>>`, one allocation per row).
    pub fn new(rows: &[Vec<NTT>], log2_b: u32, l: u32, k: u32, device: i32) -> Result<Self, CudaCommitError> {
        let (kappa, n) = (rows.len(), rows.first().map_or(0, |r| r.len()));
        let mut h = std::ptr::null_mut();
        check(unsafe { lat_ajtai_create(&mut h, kappa as u32, n as u64, log2_b, l, k, LAT_REPR_MONTGOMERY, device) }, 0, 0)?;
        let s = Self { h, kappa, n, k: k as usize };
        for (i, row) in rows.iter().enumerate() {
            check(unsafe { lat_ajtai_upload_rows(s.h, i as u32, 1, limbs(row), n as u64) }, 0, 0)?;
        }
        Ok(s)
    }
    pub fn kappa(&self) -> usize { self.kappa }
    pub fn width(&self) -> usize { self.n }

    /// commit / commit_ntt (commitment_scheme.rs:63-80,101-103)
    pub fn commit_ntt(&self, f: &[NTT]) -> Result<Vec<NTT>, CudaCommitError> {
        let mut cm = vec![NTT::default(); self.kappa];
        check(unsafe { lat_ajtai_commit_ntt(self.h, limbs(f), f.len() as u64, limbs_mut(&mut cm)) }, f.len(), self.n)?;
        Ok(cm)
    }
    /// Witness::from_w_ccs + Witness::commit fused (arith.rs:230-248,357-362; main.rs:357-363).
    /// Returns (f_coeff, f, cm).
    pub fn witness_from_w_ccs(&self, w_ccs: &[NTT]) -> Result<(Vec<Coeff>, Vec<NTT>, Vec<NTT>), CudaCommitError> {
        let mut f_coeff = vec![Coeff::default(); self.n];
        let mut f = vec![NTT::default(); self.n];
        let mut cm = vec![NTT::default(); self.kappa];
        check(
            unsafe { lat_ajtai_witness_from_w_ccs(self.h, limbs(w_ccs), w_ccs.len() as u64, limbs_mut(&mut f_coeff), limbs_mut(&mut f), limbs_mut(&mut cm)) },
            w_ccs.len() * (self.n / w_ccs.len().max(1)),
            self.n,
        )?;
        Ok((f_coeff, f, cm))
    }
    /// decompose_witness + commit_witnesses (nifs/decomposition.rs:162-201): K planes (coefficient and CRT form)
    /// and the K commitments, y_0 by homomorphism.
    pub fn decompose_commit(&self, f_coeff: &[Coeff], cm: &[NTT]) -> Result<(Vec<Vec<Coeff>>, Vec<Vec<NTT>>, Vec<Vec<NTT>>), CudaCommitError> {
        let (k, n, kappa) = (self.k, self.n, self.kappa);
        let mut pc = vec![Coeff::default(); k * n];
        let mut pf = vec![NTT::default(); k * n];
        let mut cms = vec![NTT::default(); k * kappa];
        check(
            unsafe { lat_ajtai_decompose_commit(self.h, limbs(f_coeff), f_coeff.len() as u64, limbs(cm), limbs_mut(&mut pc), limbs_mut(&mut pf), limbs_mut(&mut cms)) },
            f_coeff.len(),
            n,
        )?;
        Ok((pc.chunks(n).map(|c| c.to_vec()).collect(), pf.chunks(n).map(|c| c.to_vec()).collect(), cms.chunks(kappa).map(|c| c.to_vec()).collect()))
    }
    /// Pipelined form of `witness_from_w_ccs` + commit for a stream of VM steps (main.rs:121-219): returns at once;
    /// `w_ccs` and the returned buffer must stay alive until `wait`.  Up to 4 steps may be in flight.
    pub fn submit_w_ccs<'a>(&'a self, w_ccs: &'a [NTT]) -> Result<PendingCommit<'a>, CudaCommitError> {
        let mut cm = vec![NTT::default(); self.kappa].into_boxed_slice();
        let mut ticket = 0u64;
        check(
            unsafe { lat_ajtai_submit_w_ccs(self.h, limbs(w_ccs), w_ccs.len() as u64, cm.as_mut_ptr() as *mut u64, &mut ticket) },
            w_ccs.len() * (self.n / w_ccs.len().max(1)),
            self.n,
        )?;
        Ok(PendingCommit { engine: self, ticket, cm, _input: std::marker::PhantomData })
    }

    /// Which side (0 = accumulator, 1 = step witness) the following `decompose_commit` calls fill; both sides' planes
    /// stay resident for `fold_witness`.
    pub fn select_side(&self, side: i32) -> Result<(), CudaCommitError> {
        check(unsafe { lat_ajtai_select_side(self.h, side) }, 0, 0)
    }
    /// LFFoldingProver::compute_f_0 (nifs/folding/utils.rs:351-376) followed by Witness::from_f's iCRT
    /// (arith.rs:275-289) over the 2K resident planes.  Returns (f_0, f_0 in coefficient form).
    pub fn fold_witness(&self, rho_s: &[NTT]) -> Result<(Vec<NTT>, Vec<Coeff>), CudaCommitError> {
        assert_eq!(rho_s.len(), 2 * self.k);
        let mut f0 = vec![NTT::default(); self.n];
        let mut f0_coeff = vec![Coeff::default(); self.n];
        check(unsafe { lat_ajtai_fold_witness(self.h, limbs(rho_s), limbs_mut(&mut f0), limbs_mut(&mut f0_coeff)) }, 0, 0)?;
        Ok((f0, f0_coeff))
    }
}

/// A commitment in flight (`CudaAjtai::submit_w_ccs`); borrows the input so that it cannot be dropped early.
pub struct PendingCommit<'a> {
    engine: &'a CudaAjtai,
    ticket: u64,
    cm: Box<[NTT]>,
    _input: std::marker::PhantomData<&'a [NTT]>,
}
impl<'a> PendingCommit<'a> {
    pub fn wait(self) -> Result<Vec<NTT>, CudaCommitError> {
        check(unsafe { lat_ajtai_wait(self.engine.h, self.ticket) }, 0, 0)?;
        Ok(self.cm.into_vec())
    }
}

impl Drop for CudaAjtai {
    fn drop(&mut self) {
        unsafe { lat_ajtai_destroy(self.h) }
    }
}

/// CRT::elementwise_crt / ICRT::elementwise_icrt (ring/src/cyclotomic_ring/crt.rs:10-49), batched on the device.
pub fn elementwise_crt(v: &[Coeff], device: i32) -> Result<Vec<NTT>, CudaCommitError> {
    let mut out = vec![NTT::default(); v.len()];
    check(unsafe { lat_ring_crt(limbs(v), v.len() as u64, limbs_mut(&mut out), device) }, 0, 0)?;
    Ok(out)
}
pub fn elementwise_icrt(v: &[NTT], device: i32) -> Result<Vec<Coeff>, CudaCommitError> {
    let mut out = vec![Coeff::default(); v.len()];
    check(unsafe { lat_ring_icrt(limbs(v), v.len() as u64, limbs_mut(&mut out), device) }, 0, 0)?;
    Ok(out)
}

/// GadgetRecompose for a CRT-form vector (arith.rs:305,330; balanced_decomposition/mod.rs:177-190).
pub fn gadget_recompose(f: &[NTT], log2_b: u32, l: u32, device: i32) -> Result<Vec<NTT>, CudaCommitError> {
    let count = f.len() / l as usize;
    let mut out = vec![NTT::default(); count];
    check(unsafe { lat_ring_gadget_recompose(limbs(f), count as u64, log2_b, l, limbs_mut(&mut out), LAT_REPR_MONTGOMERY, device) }, 0, 0)?;
    Ok(out)
}

#[allow(dead_code)]
fn _unused(_: *mut c_void) {}
