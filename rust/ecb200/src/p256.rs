//! Batch methods on the `p256` types of the reference (`p256/src/lib.rs`, `p256/src/arithmetic.rs`).

use crate::curves::{self, GpuCurve};
use crate::{Engine, Error};
use ecb200_sys as sys;
use p256::{AffinePoint, NistP256, Scalar};

impl GpuCurve for NistP256 {
    const ID: i32 = sys::ECB200_P256;
    const FB: usize = 32;
}

/// `ProjectivePoint::mul_by_generator` over a slice.
pub fn mul_by_generator_batch(eng: &Engine, ks: &[Scalar], secret: bool) -> Result<Vec<AffinePoint>, Error> {
    curves::mul_by_generator_batch::<NistP256>(eng, ks, secret)
}

/// `&P * &k` + `batch_normalize` over a slice.
pub fn mul_batch(eng: &Engine, terms: &[(AffinePoint, Scalar)], secret: bool) -> Result<Vec<AffinePoint>, Error> {
    curves::mul_batch::<NistP256>(eng, terms, secret)
}

/// `LinearCombination::lincomb(&x, &k, &y, &l)` per row.
pub fn lincomb_batch(eng: &Engine, rows: &[(AffinePoint, Scalar, AffinePoint, Scalar)], secret: bool) -> Result<Vec<AffinePoint>, Error> {
    curves::lincomb_batch::<NistP256>(eng, rows, secret)
}

/// `LinearCombinationExt::lincomb_ext` (one point from many terms).
pub fn lincomb_ext(eng: &Engine, terms: &[(AffinePoint, Scalar)], secret: bool) -> Result<AffinePoint, Error> {
    curves::lincomb_ext::<NistP256>(eng, terms, secret)
}

/// `VerifyingKey::verify_prehash` over slices (`p256/src/ecdsa.rs`; no low-s rule on this curve).
pub fn verify_prehash_batch(eng: &Engine, keys: &[p256::ecdsa::VerifyingKey], prehashes: &[&[u8]], sigs: &[p256::ecdsa::Signature]) -> Result<Vec<Result<(), signature::Error>>, Error> {
    curves::verify_prehash_batch::<NistP256>(eng, keys, prehashes, sigs)
}
