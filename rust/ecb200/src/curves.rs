//! Curve-generic batch operations over the `elliptic-curve` 0.13 trait surface.  The per-curve modules instantiate
//! these with the curve id of `include/ecb200.h` and add what is curve specific (low-s rule, Schnorr, SM2DSA).

use crate::{Engine, Error};
use ecb200_sys as sys;
use ecdsa::{hazmat::bits2field, PrimeCurve, Signature, SignatureSize, VerifyingKey};
use elliptic_curve::{
    generic_array::ArrayLength,
    point::AffineCoordinates,
    sec1::{EncodedPoint, FromEncodedPoint, ModulusSize, ToEncodedPoint},
    AffinePoint, CurveArithmetic, FieldBytesSize, PrimeField, ProjectivePoint, Scalar,
};

/// A curve the engine implements: the id of `ecb200_curve` and the field size.
pub trait GpuCurve: CurveArithmetic + PrimeCurve {
    const ID: i32;
    const FB: usize;
}

fn scalars_be<C: GpuCurve>(ks: &[Scalar<C>]) -> Vec<u8> {
    let mut out = Vec::with_capacity(C::FB * ks.len());
    for k in ks {
        out.extend_from_slice(k.to_repr().as_ref()); // PrimeField::to_repr: FB bytes, big-endian
    }
    out
}

fn affine_xy<C>(p: &AffinePoint<C>, out: &mut Vec<u8>, inf: &mut Vec<u8>)
where
    C: GpuCurve,
    AffinePoint<C>: ToEncodedPoint<C>,
    FieldBytesSize<C>: ModulusSize,
{
    let e = p.to_encoded_point(false); // 04 || x || y, or 00 for the identity
    match (e.x(), e.y()) {
        (Some(x), Some(y)) => {
            out.extend_from_slice(x);
            out.extend_from_slice(y);
            inf.push(0);
        }
        _ => {
            out.resize(out.len() + 2 * C::FB, 0);
            inf.push(1);
        }
    }
}

fn slots_to_affine<C>(slots: &[u8], n: usize) -> Vec<AffinePoint<C>>
where
    C: GpuCurve,
    AffinePoint<C>: FromEncodedPoint<C> + Default,
    FieldBytesSize<C>: ModulusSize,
{
    let slot = 1 + 2 * C::FB;
    (0..n)
        .map(|i| {
            let s = &slots[i * slot..(i + 1) * slot];
            if s[0] == 0 {
                return AffinePoint::<C>::default(); // AffinePoint::IDENTITY
            }
            let e = EncodedPoint::<C>::from_bytes(s).expect("engine wrote a malformed SEC1 slot");
            Option::from(AffinePoint::<C>::from_encoded_point(&e)).expect("engine returned an off-curve point")
        })
        .collect()
}

/// `ks.iter().map(|k| (ProjectivePoint::mul_by_generator(k)).to_affine())` in one call
/// (`k256/src/arithmetic/mul.rs:415-440`, `primeorder/src/projective.rs:422-431`).  `secret = true` selects the
/// constant-time kernels (fixed windows, full table scans) - the signing / key-generation shape.
pub fn mul_by_generator_batch<C>(eng: &Engine, ks: &[Scalar<C>], secret: bool) -> Result<Vec<AffinePoint<C>>, Error>
where
    C: GpuCurve,
    AffinePoint<C>: FromEncodedPoint<C> + Default,
    FieldBytesSize<C>: ModulusSize,
{
    let kb = scalars_be::<C>(ks);
    let mut out = vec![0u8; (1 + 2 * C::FB) * ks.len()];
    let flags = sys::ECB200_FLAG_UNCOMPRESSED | if secret { sys::ECB200_FLAG_CT } else { 0 };
    eng.check(unsafe { sys::ecb200_mul_gen(eng.raw(), C::ID, ks.len(), kb.as_ptr(), out.as_mut_ptr(), flags) })?;
    Ok(slots_to_affine::<C>(&out, ks.len()))
}

/// `terms.iter().map(|(p, k)| (p * k).to_affine())` in one call (`k256/src/arithmetic/mul.rs:443-481`,
/// `primeorder/src/projective.rs:106-150` + `batch_normalize`).  Points cross the boundary in affine form: a crate
/// outside the fork cannot read (X, Y, Z); the in-fork variant passes them with `ECB200_FLAG_PROJ`.
pub fn mul_batch<C>(eng: &Engine, terms: &[(AffinePoint<C>, Scalar<C>)], secret: bool) -> Result<Vec<AffinePoint<C>>, Error>
where
    C: GpuCurve,
    AffinePoint<C>: FromEncodedPoint<C> + ToEncodedPoint<C> + Default,
    FieldBytesSize<C>: ModulusSize,
{
    let n = terms.len();
    let (mut pts, mut inf) = (Vec::with_capacity(2 * C::FB * n), Vec::with_capacity(n));
    for (p, _) in terms {
        affine_xy::<C>(p, &mut pts, &mut inf);
    }
    let ks: Vec<Scalar<C>> = terms.iter().map(|(_, k)| *k).collect();
    let kb = scalars_be::<C>(&ks);
    let mut out = vec![0u8; (1 + 2 * C::FB) * n];
    let flags = sys::ECB200_FLAG_UNCOMPRESSED | if secret { sys::ECB200_FLAG_CT } else { 0 };
    eng.check(unsafe {
        sys::ecb200_mul_var(eng.raw(), C::ID, n, pts.as_ptr(), inf.as_ptr(), kb.as_ptr(), out.as_mut_ptr(), core::ptr::null_mut(), flags)
    })?;
    Ok(slots_to_affine::<C>(&out, n))
}

/// `LinearCombination::lincomb(&x, &k, &y, &l)` per row over slices (`k256/src/arithmetic/mul.rs:313-323`,
/// `primeorder/src/projective.rs:415-420`): `out[i] = x_i * k_i + y_i * l_i`.
pub fn lincomb_batch<C>(eng: &Engine, rows: &[(AffinePoint<C>, Scalar<C>, AffinePoint<C>, Scalar<C>)], secret: bool) -> Result<Vec<AffinePoint<C>>, Error>
where
    C: GpuCurve,
    AffinePoint<C>: FromEncodedPoint<C> + ToEncodedPoint<C> + Default,
    FieldBytesSize<C>: ModulusSize,
{
    let n = rows.len();
    let (mut p1, mut p2, mut i1, mut i2) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
    for (x, _, y, _) in rows {
        affine_xy::<C>(x, &mut p1, &mut i1);
        affine_xy::<C>(y, &mut p2, &mut i2);
    }
    // an identity operand contributes nothing: feed the generator with a zero scalar instead (the C ABI takes affine points)
    let mut k1 = scalars_be::<C>(&rows.iter().map(|r| r.1).collect::<Vec<_>>());
    let mut k2 = scalars_be::<C>(&rows.iter().map(|r| r.3).collect::<Vec<_>>());
    let g = AffinePoint::<C>::from(ProjectivePoint::<C>::generator());
    let (mut gxy, mut ginf) = (Vec::new(), Vec::new());
    affine_xy::<C>(&g, &mut gxy, &mut ginf);
    for i in 0..n {
        if i1[i] != 0 {
            p1[2 * C::FB * i..2 * C::FB * (i + 1)].copy_from_slice(&gxy);
            k1[C::FB * i..C::FB * (i + 1)].fill(0);
        }
        if i2[i] != 0 {
            p2[2 * C::FB * i..2 * C::FB * (i + 1)].copy_from_slice(&gxy);
            k2[C::FB * i..C::FB * (i + 1)].fill(0);
        }
    }
    let mut out = vec![0u8; (1 + 2 * C::FB) * n];
    let flags = sys::ECB200_FLAG_UNCOMPRESSED | if secret { sys::ECB200_FLAG_CT } else { 0 };
    eng.check(unsafe {
        sys::ecb200_lincomb2(eng.raw(), C::ID, n, p1.as_ptr(), k1.as_ptr(), p2.as_ptr(), k2.as_ptr(), out.as_mut_ptr(), core::ptr::null_mut(), flags)
    })?;
    Ok(slots_to_affine::<C>(&out, n))
}

/// `LinearCombinationExt::lincomb_ext(&[(P, k)])` -> ONE point (`k256/src/arithmetic/mul.rs:326-340`).
pub fn lincomb_ext<C>(eng: &Engine, terms: &[(AffinePoint<C>, Scalar<C>)], secret: bool) -> Result<AffinePoint<C>, Error>
where
    C: GpuCurve,
    AffinePoint<C>: FromEncodedPoint<C> + ToEncodedPoint<C> + Default,
    FieldBytesSize<C>: ModulusSize,
{
    let (mut pts, mut inf) = (Vec::new(), Vec::new());
    let mut ks = Vec::new();
    for (p, k) in terms {
        let before = inf.len();
        affine_xy::<C>(p, &mut pts, &mut inf);
        if inf[before] != 0 {
            pts.truncate(pts.len() - 2 * C::FB); // the identity adds nothing to the sum
            continue;
        }
        ks.push(*k);
    }
    let kb = scalars_be::<C>(&ks);
    let mut out = vec![0u8; 1 + 2 * C::FB];
    let flags = sys::ECB200_FLAG_UNCOMPRESSED | if secret { sys::ECB200_FLAG_CT } else { 0 };
    eng.check(unsafe { sys::ecb200_lincomb(eng.raw(), C::ID, ks.len(), pts.as_ptr(), kb.as_ptr(), out.as_mut_ptr(), flags, 0) })?;
    Ok(slots_to_affine::<C>(&out, 1).remove(0))
}

/// `for ((vk, prehash), sig) in ... { vk.verify_prehash(prehash, sig) }` in one call: `PrehashVerifier::verify_prehash`
/// -> `hazmat::bits2field` (host, byte shuffling) -> `VerifyPrimitive::verify_prehashed` on the device
/// (`k256/src/ecdsa.rs:200-209` incl. the low-s rule, `p256/src/ecdsa.rs:71-75`; body in ecdsa 0.16.9 hazmat.rs).
pub fn verify_prehash_batch<C>(eng: &Engine, keys: &[VerifyingKey<C>], prehashes: &[&[u8]], sigs: &[Signature<C>]) -> Result<Vec<Result<(), signature::Error>>, Error>
where
    C: GpuCurve,
    AffinePoint<C>: FromEncodedPoint<C> + ToEncodedPoint<C> + AffineCoordinates,
    FieldBytesSize<C>: ModulusSize,
    SignatureSize<C>: ArrayLength<u8>,
{
    let n = keys.len();
    assert!(prehashes.len() == n && sigs.len() == n, "slices must have the same length");
    let (mut q, mut z, mut rs) = (vec![0u8; 2 * C::FB * n], vec![0u8; C::FB * n], vec![0u8; 2 * C::FB * n]);
    let mut pre_ok = vec![true; n];
    for i in 0..n {
        let e = keys[i].as_affine().to_encoded_point(false);
        q[2 * C::FB * i..2 * C::FB * i + C::FB].copy_from_slice(e.x().expect("a VerifyingKey is never the identity"));
        q[2 * C::FB * i + C::FB..2 * C::FB * (i + 1)].copy_from_slice(e.y().unwrap());
        match bits2field::<C>(prehashes[i]) {
            Ok(f) => z[C::FB * i..C::FB * (i + 1)].copy_from_slice(&f),
            Err(_) => pre_ok[i] = false, // prehash shorter than FB / 2 bytes
        }
        let (r, s) = sigs[i].split_bytes();
        rs[2 * C::FB * i..2 * C::FB * i + C::FB].copy_from_slice(&r);
        rs[2 * C::FB * i + C::FB..2 * C::FB * (i + 1)].copy_from_slice(&s);
    }
    let mut ok = vec![0u8; n];
    eng.check(unsafe { sys::ecb200_ecdsa_verify(eng.raw(), C::ID, n, q.as_ptr(), z.as_ptr(), rs.as_ptr(), ok.as_mut_ptr()) })?;
    Ok((0..n).map(|i| if pre_ok[i] && ok[i] == 1 { Ok(()) } else { Err(signature::Error::new()) }).collect())
}

/// `VerifyingKey::from_sec1_bytes` + `verify_prehash` with the keys decoded (decompressed) on the device.
pub fn verify_prehash_sec1_batch<C: GpuCurve>(eng: &Engine, keys: &[u8], key_stride: usize, z: &[u8], rs: &[u8]) -> Result<Vec<bool>, Error> {
    let n = z.len() / C::FB;
    assert!(keys.len() == n * key_stride && rs.len() == 2 * C::FB * n);
    let mut ok = vec![0u8; n];
    eng.check(unsafe { sys::ecb200_ecdsa_verify_sec1(eng.raw(), C::ID, n, keys.as_ptr(), key_stride, z.as_ptr(), rs.as_ptr(), ok.as_mut_ptr()) })?;
    Ok(ok.into_iter().map(|b| b == 1).collect())
}
