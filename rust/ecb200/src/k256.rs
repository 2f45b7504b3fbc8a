//! Batch methods on the `k256` types of the reference (`k256/src/lib.rs`, `k256/src/arithmetic.rs`).

use crate::curves::{self, GpuCurve};
use crate::{Engine, Error};
use ecb200_sys as sys;
use k256::{AffinePoint, Secp256k1, Scalar};

impl GpuCurve for Secp256k1 {
    const ID: i32 = sys::ECB200_K256;
    const FB: usize = 32;
}

/// `ProjectivePoint::mul_by_generator` over a slice.
pub fn mul_by_generator_batch(eng: &Engine, ks: &[Scalar], secret: bool) -> Result<Vec<AffinePoint>, Error> {
    curves::mul_by_generator_batch::<Secp256k1>(eng, ks, secret)
}

/// `&P * &k` + `batch_normalize` over a slice.
pub fn mul_batch(eng: &Engine, terms: &[(AffinePoint, Scalar)], secret: bool) -> Result<Vec<AffinePoint>, Error> {
    curves::mul_batch::<Secp256k1>(eng, terms, secret)
}

/// `LinearCombination::lincomb(&x, &k, &y, &l)` per row.
pub fn lincomb_batch(eng: &Engine, rows: &[(AffinePoint, Scalar, AffinePoint, Scalar)], secret: bool) -> Result<Vec<AffinePoint>, Error> {
    curves::lincomb_batch::<Secp256k1>(eng, rows, secret)
}

/// `LinearCombinationExt::lincomb_ext` (one point from many terms).
pub fn lincomb_ext(eng: &Engine, terms: &[(AffinePoint, Scalar)], secret: bool) -> Result<AffinePoint, Error> {
    curves::lincomb_ext::<Secp256k1>(eng, terms, secret)
}

/// `VerifyingKey::verify_prehash` over slices (`k256/src/ecdsa.rs:200-209`; high-s signatures are rejected on the device).
pub fn verify_prehash_batch(eng: &Engine, keys: &[k256::ecdsa::VerifyingKey], prehashes: &[&[u8]], sigs: &[k256::ecdsa::Signature]) -> Result<Vec<Result<(), signature::Error>>, Error> {
    curves::verify_prehash_batch::<Secp256k1>(eng, keys, prehashes, sigs)
}

/// `VerifyingKey::recover_from_prehash` over slices (`k256/src/ecdsa.rs:113-140,278-343`): SEC1 slots of the recovered keys.
pub fn recover_from_prehash_batch(eng: &Engine, z: &[u8], rs: &[u8], recid: &[u8]) -> Result<(Vec<u8>, Vec<bool>), Error> {
    let n = recid.len();
    assert!(z.len() == 32 * n && rs.len() == 64 * n);
    let (mut keys, mut ok) = (vec![0u8; 33 * n], vec![0u8; n]);
    eng.check(unsafe { sys::ecb200_ecdsa_recover(eng.raw(), sys::ECB200_K256, n, z.as_ptr(), rs.as_ptr(), recid.as_ptr(), keys.as_mut_ptr(), ok.as_mut_ptr(), 0) })?;
    Ok((keys, ok.into_iter().map(|b| b == 1).collect()))
}

/// BIP340 `schnorr::VerifyingKey::verify_prehash` after the tagged challenge hash (`k256/src/schnorr/verifying.rs:63-89`).
pub fn schnorr_verify_batch(eng: &Engine, pk: &[u8], e: &[u8], sig: &[u8]) -> Result<Vec<bool>, Error> {
    let n = pk.len() / 32;
    assert!(e.len() == 32 * n && sig.len() == 64 * n);
    let mut ok = vec![0u8; n];
    eng.check(unsafe { sys::ecb200_schnorr_verify(eng.raw(), n, pk.as_ptr(), e.as_ptr(), sig.as_ptr(), ok.as_mut_ptr()) })?;
    Ok(ok.into_iter().map(|b| b == 1).collect())
}

/// `SignPrimitive::try_sign_prehashed` over slices with caller-supplied nonces (`k256/src/ecdsa.rs:181-198`):
/// (r || s, recovery id, ok) per row; constant-time kernels, staged secrets wiped by the library after the call.
pub fn try_sign_prehashed_batch(eng: &Engine, d: &[u8], k: &[u8], z: &[u8]) -> Result<(Vec<u8>, Vec<u8>, Vec<bool>), Error> {
    let n = z.len() / 32;
    assert!(d.len() == 32 * n && k.len() == 32 * n);
    let (mut rs, mut recid, mut ok) = (vec![0u8; 64 * n], vec![0u8; n], vec![0u8; n]);
    eng.check(unsafe { sys::ecb200_ecdsa_sign(eng.raw(), sys::ECB200_K256, n, d.as_ptr(), k.as_ptr(), z.as_ptr(), rs.as_mut_ptr(), recid.as_mut_ptr(), ok.as_mut_ptr()) })?;
    Ok((rs, recid, ok.into_iter().map(|b| b == 1).collect()))
}
