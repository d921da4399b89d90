//! `EvaluationDomain<Fr>` on the device ([UPSTREAM] halo2_proofs/src/poly/domain.rs). patches/domain.patch adds
//!     #[cfg(feature = "b200")] gpu: std::sync::OnceLock<Option<crate::b200::GpuDomain>>,
//! to the struct (initialised empty in `new`) and makes the four methods below try the device first. The kernels fold the
//! zeta-coset scaling, the zero padding, the 1/n factors and the truncation into the first load / last store of the NTT,
//! so each method is exactly one library call.
//!
//! Only the `F = Fr` instantiation takes the device path (TypeId check, as in arithmetic_b200.rs).

use crate::b200::GpuDomain;
use crate::poly::{Coeff, EvaluationDomain, ExtendedLagrangeCoeff, LagrangeCoeff, Polynomial};
use halo2curves::bn256::Fr;

/// below this k the CPU is faster than the PCIe round trip of the host-buffer entry points
const MIN_K: u32 = 14;

impl EvaluationDomain<Fr> {
    fn gpu(&self) -> Option<&GpuDomain> {
        if self.k() < MIN_K {
            return None;
        }
        // j = quotient_poly_degree + 1: spb_domain_new derives extended_k exactly as EvaluationDomain::new does
        self.gpu.get_or_init(|| GpuDomain::new(self.get_quotient_poly_degree() as u32 + 1, self.k())).as_ref()
    }

    pub(crate) fn lagrange_to_coeff_b200(&self, a: &mut Polynomial<Fr, LagrangeCoeff>) -> bool {
        self.gpu().and_then(|g| g.lagrange_to_coeff(&mut a.values)).is_some()
    }
    pub(crate) fn coeff_to_extended_b200(&self, a: &Polynomial<Fr, Coeff>, out: &mut Polynomial<Fr, ExtendedLagrangeCoeff>) -> bool {
        self.gpu().and_then(|g| g.coeff_to_extended(&a.values, &mut out.values)).is_some()
    }
    pub(crate) fn extended_to_coeff_b200(&self, a: &Polynomial<Fr, ExtendedLagrangeCoeff>, out: &mut Vec<Fr>) -> bool {
        self.gpu().and_then(|g| g.extended_to_coeff(&a.values, out)).is_some()
    }
    pub(crate) fn divide_by_vanishing_poly_b200(&self, a: &mut Polynomial<Fr, ExtendedLagrangeCoeff>) -> bool {
        self.gpu().and_then(|g| g.divide_by_vanishing_poly(&mut a.values)).is_some()
    }
}
