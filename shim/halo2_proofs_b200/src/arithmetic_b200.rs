//! Front doors of `halo2_proofs::arithmetic::{best_multiexp, best_fft}` ([UPSTREAM] halo2_proofs/src/arithmetic.rs; reached
//! from Spectre at lightclient-circuits/src/util/circuit.rs:131,158,177,211,263 through snark_verifier_sdk).
//!
//! How to apply (patches/arithmetic.patch does exactly this): in the fork's `src/arithmetic.rs`
//!   1. rename the upstream bodies `best_multiexp` -> `best_multiexp_cpu`, `best_fft` -> `best_fft_cpu` (unchanged);
//!   2. add `#[cfg(feature = "b200")] mod arithmetic_b200; ` next to the other `mod` lines of lib.rs and
//!      `pub use` nothing -- the two functions below take the upstream names and signatures.
//! Generic code keeps compiling: the device path is taken only when the concrete types are BN254's (checked with
//! `TypeId`, a branch the optimiser removes after monomorphisation), everything else falls through to the CPU body.

use crate::arithmetic::{best_fft_cpu, best_multiexp_cpu, CurveAffine, FftGroup};
use crate::b200;
use ff::Field;
use halo2curves::bn256::{Fr, G1Affine, G1};
use std::any::TypeId;

/// below these sizes the PCIe round trip costs more than the CPU needs (measured crossover)
const MSM_MIN: usize = 1 << 12;
const FFT_MIN_LOG: u32 = 14;

pub fn best_multiexp<C: CurveAffine>(coeffs: &[C::Scalar], bases: &[C]) -> C::Curve {
    assert_eq!(coeffs.len(), bases.len());
    if coeffs.len() >= MSM_MIN && TypeId::of::<C>() == TypeId::of::<G1Affine>() {
        // SAFETY: C == G1Affine, C::Scalar == Fr, C::Curve == G1 (TypeId check above); same layout, same lifetime.
        let (s, b) = unsafe {
            (
                std::slice::from_raw_parts(coeffs.as_ptr() as *const Fr, coeffs.len()),
                std::slice::from_raw_parts(bases.as_ptr() as *const G1Affine, bases.len()),
            )
        };
        if let Some(r) = b200::msm_raw(s, b) {
            return unsafe { std::mem::transmute_copy::<G1, C::Curve>(&r) };
        }
    }
    best_multiexp_cpu(coeffs, bases)
}

pub fn best_fft<Scalar: Field, G: FftGroup<Scalar>>(a: &mut [G], omega: Scalar, log_n: u32) {
    if log_n >= FFT_MIN_LOG && TypeId::of::<G>() == TypeId::of::<Fr>() && TypeId::of::<Scalar>() == TypeId::of::<Fr>() {
        // SAFETY: G == Scalar == Fr.
        let v = unsafe { std::slice::from_raw_parts_mut(a.as_mut_ptr() as *mut Fr, a.len()) };
        let w = unsafe { std::mem::transmute_copy::<Scalar, Fr>(&omega) };
        if b200::ntt(v, w, log_n).is_some() {
            return;
        }
    }
    best_fft_cpu(a, omega, log_n)
}
