//! Batch entry points for the hot path of the RustCrypto `elliptic-curves` workspace (risc0 fork), backed by the
//! B200 engine behind `include/ecb200.h`.
//!
//! The reference exposes per-element operations - `ProjectivePoint::mul_by_generator(&k)`, `&p * &k`,
//! `LinearCombination::lincomb`, `BatchNormalize::batch_normalize`, `VerifyingKey::verify_prehash` - and callers loop
//! over slices (`k256/benches/ecdsa.rs:57-63`, `k256/benches/scalar.rs:59-65`).  This crate keeps the reference's types
//! and adds ONE call per slice; serialisation goes through the public accessors (`Scalar::to_repr`, `ToEncodedPoint`,
//! `Signature::split_bytes`), because point coordinates are private outside the curve crates
//! (`k256/src/arithmetic/projective.rs:38-42`, `primeorder/src/projective.rs:37-41`).
//!
//! SOURCE ONLY in this repository: the build image has no Rust toolchain, so these files are not compiled or tested
//! here; `tests/test_rust_ffi.py` checks mechanically that `ecb200-sys` declares every symbol of the header with the
//! right arity and types and that every `ecb200_sys::` call made here names an existing symbol with the right
//! argument count.  The C++ mirror `include/ecb200.hpp` is the compiled, tested host side.
//!
//! There is no CPU fallback: [`Engine::new`] fails when no B200 is usable.

use core::ffi::CStr;
use ecb200_sys as sys;

pub mod curves;
#[cfg(feature = "k256")]
pub mod k256;
#[cfg(feature = "p256")]
pub mod p256;
#[cfg(feature = "p384")]
pub mod p384;
#[cfg(feature = "sm2")]
pub mod sm2;

/// Engine failure (no CUDA device, launch or copy error, bad argument).  Per-element failures are data
/// (`Result<(), signature::Error>` per row), never this.
#[derive(Debug, Clone, PartialEq, Eq)]
pub struct Error {
    pub status: i32,
    pub message: String,
}

impl core::fmt::Display for Error {
    fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
        write!(f, "ecb200 status {}: {}", self.status, self.message)
    }
}
impl std::error::Error for Error {}

/// One engine context: one B200 ([`Engine::new`]) or several GPUs of the box ([`Engine::new_multi`]), in which case
/// every batch call shards its rows by contiguous index range over the devices (no collective on the data path).
/// Thread-compatible (`Send`, not `Sync`): one call at a time per context.
pub struct Engine {
    ctx: *mut sys::ecb200_ctx,
}

unsafe impl Send for Engine {}

impl Engine {
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut ctx = core::ptr::null_mut();
        let rc = unsafe { sys::ecb200_init(device, &mut ctx) };
        if rc != 0 || ctx.is_null() {
            return Err(Error { status: rc, message: "ecb200_init failed: no usable sm_100 device (there is no CPU fallback)".into() });
        }
        Ok(Self { ctx })
    }

    /// `devices = &[]` means every visible device.
    pub fn new_multi(devices: &[i32]) -> Result<Self, Error> {
        let mut ctx = core::ptr::null_mut();
        let ptr = if devices.is_empty() { core::ptr::null() } else { devices.as_ptr() };
        let rc = unsafe { sys::ecb200_init_multi(devices.len() as i32, ptr, &mut ctx) };
        if rc != 0 || ctx.is_null() {
            return Err(Error { status: rc, message: "ecb200_init_multi failed".into() });
        }
        Ok(Self { ctx })
    }

    pub fn device_count(&self) -> usize {
        unsafe { sys::ecb200_device_count(self.ctx) as usize }
    }

    pub fn launch_count(&self) -> u64 {
        unsafe { sys::ecb200_launch_count(self.ctx) }
    }

    pub(crate) fn raw(&self) -> *mut sys::ecb200_ctx {
        self.ctx
    }

    pub(crate) fn check(&self, rc: i32) -> Result<(), Error> {
        if rc == 0 {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(sys::ecb200_last_error(self.ctx)) }.to_string_lossy().into_owned();
        Err(Error { status: rc, message: msg })
    }
}

impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { sys::ecb200_destroy(self.ctx) }
    }
}
