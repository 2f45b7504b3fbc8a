//! Reference arm with the REAL crates: `VerifyingKey::verify_prehash` (k256/src/ecdsa.rs:200-209, p256/src/ecdsa.rs:71-75)
//! over the rows of one bench.py batch, `rayon` over every host core, one JSON line in bench.py's format.
//!
//! Not compiled in this repository's image (no Rust toolchain) - bench.py times the C++ port of the same algorithms
//! (oracle/ecport.cpp) instead and says so (`cpu_baseline.kind: "port"`).  Input: a directory written by
//! `python tools/dump_batch.py <curve> <log2 rows> <dir>` with q.bin (n x 64, x||y), z.bin (n x 32), rs.bin (n x 64, r||s)
//! and expected.bin (n bytes, the mask implied by the construction of the batch - the same rows, seed and corruption
//! pattern our arm verifies).  Keys are parsed before the clock starts (the GPU arm validates them inside its timed region).

use rayon::prelude::*;
use std::{env, fs, path::Path, time::Instant};

macro_rules! verify_rows {
    ($krate:ident, $q:expr, $z:expr, $rs:expr) => {{
        use $krate::ecdsa::{signature::hazmat::PrehashVerifier, Signature, VerifyingKey};
        use $krate::{EncodedPoint, FieldBytes};
        let n = $z.len() / 32;
        let keys: Vec<Option<VerifyingKey>> = (0..n)
            .into_par_iter()
            .map(|i| {
                let x = FieldBytes::clone_from_slice(&$q[64 * i..64 * i + 32]);
                let y = FieldBytes::clone_from_slice(&$q[64 * i + 32..64 * i + 64]);
                VerifyingKey::from_encoded_point(&EncodedPoint::from_affine_coordinates(&x, &y, false)).ok()
            })
            .collect();
        let t = Instant::now();
        let ok: Vec<u8> = (0..n)
            .into_par_iter()
            .map(|i| match (&keys[i], Signature::from_slice(&$rs[64 * i..64 * i + 64])) {
                (Some(vk), Ok(sig)) => vk.verify_prehash(&$z[32 * i..32 * i + 32], &sig).is_ok() as u8,
                _ => 0, // out-of-range r / s never reach the arithmetic (ecdsa 0.16.9 Signature construction)
            })
            .collect();
        (ok, t.elapsed().as_secs_f64())
    }};
}

fn main() {
    let args: Vec<String> = env::args().collect();
    let dir = Path::new(args.get(1).expect("usage: ecb200-refbench <dir> [k256|p256]"));
    let curve = args.get(2).map(String::as_str).unwrap_or("k256");
    let rd = |f: &str| fs::read(dir.join(f)).unwrap_or_else(|e| panic!("{f}: {e}"));
    let (q, z, rs, expected) = (rd("q.bin"), rd("z.bin"), rd("rs.bin"), rd("expected.bin"));
    let n = z.len() / 32;
    assert!(q.len() == 64 * n && rs.len() == 64 * n && expected.len() == n, "inconsistent batch files");
    let (ok, secs) = match curve {
        "k256" => verify_rows!(k256, q, z, rs),
        "p256" => verify_rows!(p256, q, z, rs),
        other => panic!("unsupported curve {other}"),
    };
    let cores = rayon::current_num_threads();
    let value = n as f64 / secs;
    println!(
        "{{\"impl\": \"reference\", \"metric\": \"{curve} ECDSA verify_prehash throughput\", \"value\": {value:.1}, \"unit\": \"verifies/s\", \
         \"higher_is_better\": true, \"cpu_baseline\": {{\"value\": {value:.1}, \"unit\": \"verifies/s\", \"cores\": {cores}, \"kind\": \"reference\", \
         \"sample\": \"{n} rows of the bench.py batch, one pass ({secs:.2} s)\", \"per_core\": {:.1}, \"matches_constructed_mask\": {}}}}}",
        value / cores as f64,
        ok == expected
    );
}
