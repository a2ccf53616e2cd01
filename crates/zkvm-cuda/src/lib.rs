//! Safe Rust wrapper over the C ABI of the B200 Ajtai commitment engine (`include/lattice_ajtai.h`).
//!
//! `latticefold` is `#![forbid(unsafe_code)]` (crates/latticefold/src/lib.rs:4), so all `unsafe` lives here and
//! `latticefold` / `zkvm` call the safe API below.  A `GoldilocksRingNTT` / `GoldilocksRingPoly` is exactly 24
//! contiguous `u64` Montgomery limbs (stark-rings ring/src/cyclotomic_ring/flatten.rs:10-17,
//! models/goldilocks/utils.rs:5-23), so slices are passed by pointer with `LAT_REPR_MONTGOMERY` and no conversion.
//!
//! Thread safety.  The C handle carries unsynchronised mutable state (work buffers, the resident witness, pipeline
//! slots, `cudaSetDevice`), so every method that touches the device takes `&mut self`: the borrow checker then gives
//! the C ABI's "one caller at a time per handle" for free.  `CudaAjtai` is `Send` (the handle may move to another
//! thread) but deliberately NOT `Sync`; code that calls `commit` from rayon iterators (`commit_witnesses` does,
//! crates/latticefold/src/nifs/decomposition.rs:184-186) must go through the batched `decompose_commit` /
//! `fold_step_begin` entry points instead, or wrap the engine in a `Mutex`.
//!
//! NOT compiled in the build environment of the engine (no cargo/rustc there): see INTEGRATION.md.
use cyclotomic_rings::rings::{GoldilocksRingNTT as NTT, GoldilocksRingPoly as Coeff};
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct LatAjtai {
    _private: [u8; 0],
}

pub const LAT_REPR_MONTGOMERY: c_int = 1;

extern "C" {
    fn lat_last_error() -> *const c_char;
    fn lat_ajtai_create(out: *mut *mut LatAjtai, kappa: u32, n: u64, log2_b: u32, l: u32, k: u32, repr: c_int, device: c_int) -> c_int;
    fn lat_ajtai_destroy(h: *mut LatAjtai);
    fn lat_ajtai_upload_rows(h: *mut LatAjtai, row0: u32, nrows: u32, rows: *const u64, row_stride: u64) -> c_int;
    fn lat_ajtai_commit_ntt(h: *mut LatAjtai, f: *const u64, f_len: u64, cm: *mut u64) -> c_int;
    fn lat_ajtai_witness_from_w_ccs(h: *mut LatAjtai, w_ccs: *const u64, w_len: u64, f_coeff: *mut u64, f: *mut u64, cm: *mut u64) -> c_int;
    fn lat_ajtai_witness_from_w_ccs_compact(h: *mut LatAjtai, w_ccs: *const u64, w_len: u64, f_coeff16: *mut i16, f: *mut u64, cm: *mut u64) -> c_int;
    fn lat_ajtai_decompose_commit(h: *mut LatAjtai, f_coeff: *const u64, n: u64, cm: *const u64, planes_coeff: *mut u64, planes_f: *mut u64, cms: *mut u64) -> c_int;
    fn lat_ajtai_submit_w_ccs(h: *mut LatAjtai, w_ccs: *const u64, w_len: u64, cm: *mut u64, ticket: *mut u64) -> c_int;
    fn lat_ajtai_wait(h: *mut LatAjtai, ticket: u64) -> c_int;
    fn lat_ajtai_select_side(h: *mut LatAjtai, side: c_int) -> c_int;
    fn lat_ajtai_fold_witness(h: *mut LatAjtai, rho: *const u64, f0: *mut u64, f0_coeff: *mut u64) -> c_int;
    fn lat_ajtai_set_accumulator(h: *mut LatAjtai, f_coeff: *const u64, n: u64, cm_acc: *const u64) -> c_int;
    fn lat_ajtai_fold_step_begin(h: *mut LatAjtai, w_ccs: *const u64, w_len: u64, cm_acc: *const u64, f_coeff16: *mut i16, cm: *mut u64, cms: *mut u64) -> c_int;
    fn lat_ajtai_fold_step_finish(h: *mut LatAjtai, rho: *const u64, f0_coeff16: *mut i16, f0: *mut u64, cm0: *mut u64, w_ccs0: *mut u64) -> c_int;
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
fn opt_limbs<T>(v: Option<&[T]>) -> *const u64 {
    v.map_or(std::ptr::null(), limbs)
}

/// Device-resident replacement of `AjtaiCommitmentScheme<GoldilocksRingNTT>`.
pub struct CudaAjtai {
    h: *mut LatAjtai,
    kappa: usize,
    n: usize,
    l: usize,
    k: usize,
}
// The handle is an owned pointer to C state that one thread at a time may use: it can move between threads, and every
// method that reaches the device needs `&mut self`.  No `Sync`: see the module docs.
unsafe impl Send for CudaAjtai {}

impl CudaAjtai {
    /// AjtaiCommitmentScheme::new(Matrix<R>) (commitment_scheme.rs:49): uploads the host's rows as they are
    /// (`Vec<Vec<R>>`, one allocation per row).
    pub fn new(rows: &[Vec<NTT>], log2_b: u32, l: u32, k: u32, device: i32) -> Result<Self, CudaCommitError> {
        let (kappa, n) = (rows.len(), rows.first().map_or(0, |r| r.len()));
        let mut h = std::ptr::null_mut();
        check(unsafe { lat_ajtai_create(&mut h, kappa as u32, n as u64, log2_b, l, k, LAT_REPR_MONTGOMERY, device) }, 0, 0)?;
        let s = Self { h, kappa, n, l: l as usize, k: k as usize };
        for (i, row) in rows.iter().enumerate() {
            check(unsafe { lat_ajtai_upload_rows(s.h, i as u32, 1, limbs(row), n as u64) }, 0, 0)?;
        }
        Ok(s)
    }
    pub fn kappa(&self) -> usize { self.kappa }
    pub fn width(&self) -> usize { self.n }

    /// commit / commit_ntt (commitment_scheme.rs:63-80,101-103)
    pub fn commit_ntt(&mut self, f: &[NTT]) -> Result<Vec<NTT>, CudaCommitError> {
        let mut cm = vec![NTT::default(); self.kappa];
        check(unsafe { lat_ajtai_commit_ntt(self.h, limbs(f), f.len() as u64, limbs_mut(&mut cm)) }, f.len(), self.n)?;
        Ok(cm)
    }
    /// Witness::from_w_ccs + Witness::commit fused (arith.rs:230-248,357-362; main.rs:357-363).
    /// Returns (f_coeff, f, cm).
    pub fn witness_from_w_ccs(&mut self, w_ccs: &[NTT]) -> Result<(Vec<Coeff>, Vec<NTT>, Vec<NTT>), CudaCommitError> {
        let mut f_coeff = vec![Coeff::default(); self.n];
        let mut f = vec![NTT::default(); self.n];
        let mut cm = vec![NTT::default(); self.kappa];
        check(
            unsafe { lat_ajtai_witness_from_w_ccs(self.h, limbs(w_ccs), w_ccs.len() as u64, limbs_mut(&mut f_coeff), limbs_mut(&mut f), limbs_mut(&mut cm)) },
            w_ccs.len() * self.l,
            self.n,
        )?;
        Ok((f_coeff, f, cm))
    }
    /// The same, fetching `f_coeff` as the int16 digits the device holds (n x 24; 4.7 MB instead of 19 MB at the zkVM's
    /// size) and no `f`: `get_fhat` (arith.rs:273-297) is a re-layout of exactly these digits, `digit as i128` into `Fq`.
    pub fn witness_from_w_ccs_compact(&mut self, w_ccs: &[NTT]) -> Result<(Vec<[i16; 24]>, Vec<NTT>), CudaCommitError> {
        let mut digits = vec![[0i16; 24]; self.n];
        let mut cm = vec![NTT::default(); self.kappa];
        check(
            unsafe { lat_ajtai_witness_from_w_ccs_compact(self.h, limbs(w_ccs), w_ccs.len() as u64, digits.as_mut_ptr() as *mut i16, std::ptr::null_mut(), limbs_mut(&mut cm)) },
            w_ccs.len() * self.l,
            self.n,
        )?;
        Ok((digits, cm))
    }
    /// decompose_witness + commit_witnesses (nifs/decomposition.rs:162-201): K planes (coefficient and CRT form)
    /// and the K commitments, y_0 by homomorphism.
    pub fn decompose_commit(&mut self, f_coeff: &[Coeff], cm: &[NTT]) -> Result<(Vec<Vec<Coeff>>, Vec<Vec<NTT>>, Vec<Vec<NTT>>), CudaCommitError> {
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
    /// Non-blocking `witness_from_w_ccs` + commit: returns at once; `w_ccs` stays borrowed until `wait` (or drop).
    /// For work that is independent of the commitment.  Consecutive IVC steps are NOT independent (step i+1's `z` is
    /// built from `fold(cm_i, w_i)`, zkvm/src/main.rs:140-156,174-182), so the folding loop uses the blocking calls.
    /// Up to 4 tickets may be in flight.
    pub fn submit_w_ccs<'a>(&'a mut self, w_ccs: &'a [NTT]) -> Result<PendingCommit<'a>, CudaCommitError> {
        let mut cm = vec![NTT::default(); self.kappa].into_boxed_slice();
        let mut ticket = 0u64;
        check(
            unsafe { lat_ajtai_submit_w_ccs(self.h, limbs(w_ccs), w_ccs.len() as u64, cm.as_mut_ptr() as *mut u64, &mut ticket) },
            w_ccs.len() * self.l,
            self.n,
        )?;
        Ok(PendingCommit { h: self.h, ticket, cm: Some(cm), _borrows: std::marker::PhantomData })
    }

    /// Which side (0 = accumulator, 1 = step witness) the following `decompose_commit` calls fill; both sides' planes
    /// stay resident for `fold_witness`.
    pub fn select_side(&mut self, side: i32) -> Result<(), CudaCommitError> {
        check(unsafe { lat_ajtai_select_side(self.h, side) }, 0, 0)
    }
    /// LFFoldingProver::compute_f_0 (nifs/folding.rs:258-268) followed by Witness::from_f's iCRT (arith.rs:299-313)
    /// over the 2K resident planes.  Returns (f_0, f_0 in coefficient form).
    pub fn fold_witness(&mut self, rho_s: &[NTT]) -> Result<(Vec<NTT>, Vec<Coeff>), CudaCommitError> {
        assert_eq!(rho_s.len(), 2 * self.k);
        let mut f0 = vec![NTT::default(); self.n];
        let mut f0_coeff = vec![Coeff::default(); self.n];
        check(unsafe { lat_ajtai_fold_witness(self.h, limbs(rho_s), limbs_mut(&mut f0), limbs_mut(&mut f0_coeff)) }, 0, 0)?;
        Ok((f0, f0_coeff))
    }

    // ---- the fold step of one IVC step as two blocking calls (zk_latticefold.rs:37-102; main.rs:174-182) -------------
    /// initialize_accumulator (main.rs:306-344): the accumulator witness's coefficients and its commitment go up once;
    /// from then on the accumulator stays on the device.
    pub fn set_accumulator(&mut self, f_coeff: &[Coeff], cm_acc: &[NTT]) -> Result<(), CudaCommitError> {
        check(unsafe { lat_ajtai_set_accumulator(self.h, limbs(f_coeff), f_coeff.len() as u64, limbs(cm_acc)) }, f_coeff.len(), self.n)
    }
    /// `commit(z)` of the step witness + `decompose_witness`/`commit_witnesses` of the accumulator (side 0) and of the
    /// step witness (side 1).  `cm_acc = None` uses the folded commitment the previous `fold_step_finish` left resident.
    /// Returns (cm_i, y_s of the accumulator, y_s of the step witness, the step witness's digits).
    #[allow(clippy::type_complexity)]
    pub fn fold_step_begin(&mut self, w_ccs: &[NTT], cm_acc: Option<&[NTT]>) -> Result<(Vec<NTT>, Vec<Vec<NTT>>, Vec<Vec<NTT>>, Vec<[i16; 24]>), CudaCommitError> {
        let (k, kappa) = (self.k, self.kappa);
        let mut digits = vec![[0i16; 24]; self.n];
        let mut cm = vec![NTT::default(); kappa];
        let mut cms = vec![NTT::default(); 2 * k * kappa];
        check(
            unsafe { lat_ajtai_fold_step_begin(self.h, limbs(w_ccs), w_ccs.len() as u64, opt_limbs(cm_acc), digits.as_mut_ptr() as *mut i16, limbs_mut(&mut cm), limbs_mut(&mut cms)) },
            w_ccs.len() * self.l,
            self.n,
        )?;
        let side = |s: usize| cms[s * k * kappa..(s + 1) * k * kappa].chunks(kappa).map(|c| c.to_vec()).collect::<Vec<_>>();
        Ok((cm, side(0), side(1), digits))
    }
    /// compute_f_0 (nifs/folding.rs:258-268) + Witness::from_f (arith.rs:299-313) + cm_0 = sum rho_i cm_i
    /// (nifs/folding/utils.rs:466-472).  Returns (cm_0, f_0's digits, w_ccs of the folded witness); f_0 itself stays on the
    /// device as the next step's accumulator.  `DigitOverflow` = the folded witness broke the norm bound 2^K.
    pub fn fold_step_finish(&mut self, rho_s: &[NTT]) -> Result<(Vec<NTT>, Vec<[i16; 24]>, Vec<NTT>), CudaCommitError> {
        assert_eq!(rho_s.len(), 2 * self.k);
        let mut digits = vec![[0i16; 24]; self.n];
        let mut cm0 = vec![NTT::default(); self.kappa];
        let mut w_ccs0 = vec![NTT::default(); self.n / self.l];
        check(
            unsafe { lat_ajtai_fold_step_finish(self.h, limbs(rho_s), digits.as_mut_ptr() as *mut i16, std::ptr::null_mut(), limbs_mut(&mut cm0), limbs_mut(&mut w_ccs0)) },
            0,
            0,
        )?;
        Ok((cm0, digits, w_ccs0))
    }
}

/// A commitment in flight (`CudaAjtai::submit_w_ccs`).  It borrows the engine mutably and the input immutably until it
/// is waited for; dropping it without `wait` (an early `?` return) still waits for the ticket, which frees the engine's
/// pipeline slot and ends the device's use of both buffers before the borrows end.
pub struct PendingCommit<'a> {
    h: *mut LatAjtai,
    ticket: u64,
    cm: Option<Box<[NTT]>>,
    _borrows: std::marker::PhantomData<(&'a mut CudaAjtai, &'a [NTT])>,
}
impl<'a> PendingCommit<'a> {
    pub fn wait(mut self) -> Result<Vec<NTT>, CudaCommitError> {
        let cm = self.cm.take().expect("waited once");
        check(unsafe { lat_ajtai_wait(self.h, self.ticket) }, 0, 0)?;
        Ok(cm.into_vec())
    }
}
impl<'a> Drop for PendingCommit<'a> {
    fn drop(&mut self) {
        if self.cm.is_some() {
            // never waited for: the device may still be reading the input and will still write `cm`
            let _ = unsafe { lat_ajtai_wait(self.h, self.ticket) };
        }
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
