//! Batch methods on the `sm2` types of the reference (`sm2/src/lib.rs`, `sm2/src/arithmetic.rs`).

use crate::curves::{self, GpuCurve};
use crate::{Engine, Error};
use ecb200_sys as sys;
use sm2::{AffinePoint, Sm2, Scalar};

impl GpuCurve for Sm2 {
    const ID: i32 = sys::ECB200_SM2;
    const FB: usize = 32;
}

/// `ProjectivePoint::mul_by_generator` over a slice.
pub fn mul_by_generator_batch(eng: &Engine, ks: &[Scalar], secret: bool) -> Result<Vec<AffinePoint>, Error> {
    curves::mul_by_generator_batch::<Sm2>(eng, ks, secret)
}

/// `&P * &k` + `batch_normalize` over a slice.
pub fn mul_batch(eng: &Engine, terms: &[(AffinePoint, Scalar)], secret: bool) -> Result<Vec<AffinePoint>, Error> {
    curves::mul_batch::<Sm2>(eng, terms, secret)
}

/// `LinearCombination::lincomb(&x, &k, &y, &l)` per row.
pub fn lincomb_batch(eng: &Engine, rows: &[(AffinePoint, Scalar, AffinePoint, Scalar)], secret: bool) -> Result<Vec<AffinePoint>, Error> {
    curves::lincomb_batch::<Sm2>(eng, rows, secret)
}

/// `LinearCombinationExt::lincomb_ext` (one point from many terms).
pub fn lincomb_ext(eng: &Engine, terms: &[(AffinePoint, Scalar)], secret: bool) -> Result<AffinePoint, Error> {
    curves::lincomb_ext::<Sm2>(eng, terms, secret)
}

/// `sm2::dsa::VerifyingKey::verify_prehash` over slices (`sm2/src/dsa/verifying.rs:130-168`): q = x || y keys,
/// e = SM3(Z_A || M) digests, rs = r || s.
pub fn sm2dsa_verify_batch(eng: &Engine, q: &[u8], e: &[u8], rs: &[u8]) -> Result<Vec<bool>, Error> {
    let n = e.len() / 32;
    assert!(q.len() == 64 * n && rs.len() == 64 * n);
    let mut ok = vec![0u8; n];
    eng.check(unsafe { sys::ecb200_sm2dsa_verify(eng.raw(), n, q.as_ptr(), e.as_ptr(), rs.as_ptr(), ok.as_mut_ptr()) })?;
    Ok(ok.into_iter().map(|b| b == 1).collect())
}
