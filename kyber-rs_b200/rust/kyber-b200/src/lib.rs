//! Batch entry points behind kyber-rs's existing trait surface (src/group.rs: `Point` :85,
//! `Scalar` :22).  NOT COMPILED in the environment this repository is built in.
//!
//! Protocol code keeps using `Group`/`Point`/`Scalar`; where it loops over `Point::mul`,
//! `PubPoly::check` or `schnorr::verify` today it can hand the whole loop to one call here and
//! gets, item by item, exactly what the scalar call would have returned (compressed encodings,
//! `SignatureError` variants) — that bit-exactness is what `tests/` checks through the same C ABI.
use kyber_b200_sys as sys;
use kyber_rs::encoding::{BinaryMarshaler, BinaryUnmarshaler, MarshallingError};
use kyber_rs::group::edwards25519::{Point as EdPoint, Scalar as EdScalar};
use kyber_rs::share::poly::{PriShare, PubPoly};
use kyber_rs::sign::error::SignatureError;

#[derive(Debug)]
pub enum GpuError {
    Arg,
    Cuda(String),
    NoMem,
    /// an input point does not decode (what `Point::unmarshal_binary` reports as MarshallingError::InvalidInput);
    /// the index of the first such item
    InvalidPoint(usize),
}

/// One CUDA device, one host thread — kyber-rs itself is single-threaded.
pub struct Gpu {
    ctx: *mut sys::kb_ctx,
}

impl Gpu {
    pub fn new(device: i32) -> Result<Self, GpuError> {
        let mut ctx = std::ptr::null_mut();
        match unsafe { sys::kb_ctx_create(device, &mut ctx) } {
            sys::KB_OK => Ok(Gpu { ctx }),
            sys::KB_ERR_NOMEM => Err(GpuError::NoMem),
            _ => Err(GpuError::Cuda("no usable CUDA device (there is no CPU fallback)".into())),
        }
    }
    fn check(&self, rc: i32) -> Result<(), GpuError> {
        match rc {
            sys::KB_OK => Ok(()),
            sys::KB_ERR_ARG => Err(GpuError::Arg),
            sys::KB_ERR_NOMEM => Err(GpuError::NoMem),
            _ => Err(GpuError::Cuda(unsafe { std::ffi::CStr::from_ptr(sys::kb_last_error(self.ctx)) }.to_string_lossy().into_owned())),
        }
    }
}
impl Drop for Gpu {
    fn drop(&mut self) {
        unsafe { sys::kb_ctx_destroy(self.ctx) }
    }
}

fn pack_scalars(s: &[EdScalar]) -> Vec<u8> {
    s.iter().flat_map(|x| x.v).collect() // Scalar.v: [u8; 32], raw (scalar.rs:24)
}
fn pack_points(p: &[EdPoint]) -> Vec<u8> {
    p.iter().flat_map(|x| x.marshal_binary().expect("32 bytes")).collect()
}
/// The in-memory form of a `Point` (ge.rs:75-83: X, Y, Z, T as 10 i32 limbs each) — what the `_limbs` / `KB_POINT_LIMBS40`
/// entry points take, so that a caller holding thousands of commitments does not pay one field inversion each for
/// `marshal_binary`.
fn pack_point_limbs(p: &[EdPoint]) -> Vec<i32> {
    p.iter().flat_map(|x| x.ge_mut_ref_limbs()).collect() // accessor to be added next to Point::ge (point.rs:24)
}
fn unpack_points(b: &[u8]) -> Vec<EdPoint> {
    b.chunks(32)
        .map(|c| {
            let mut p = EdPoint::default();
            p.unmarshal_binary(c).expect("library output is a canonical encoding");
            p
        })
        .collect()
}

/// Batch companion of `group::Point::mul` (group.rs:139; point.rs:207-225).
pub trait BatchPoint: Sized {
    /// `points = None` multiplies the standard base (ge_scalar_mult_base, ge.rs:442).
    fn mul_batch(gpu: &Gpu, scalars: &[EdScalar], points: Option<&[Self]>) -> Result<Vec<Self>, GpuError>;
}
impl BatchPoint for EdPoint {
    fn mul_batch(gpu: &Gpu, scalars: &[EdScalar], points: Option<&[EdPoint]>) -> Result<Vec<EdPoint>, GpuError> {
        let n = scalars.len();
        let s = pack_scalars(scalars);
        let mut out = vec![0u8; 32 * n];
        match points {
            None => gpu.check(unsafe { sys::kb_point_mul_base_batch(gpu.ctx, n, s.as_ptr(), out.as_mut_ptr(), 0) })?,
            Some(ps) => {
                assert_eq!(ps.len(), n);
                let p = pack_points(ps);
                let mut st = vec![0u8; n];
                gpu.check(unsafe { sys::kb_point_mul_batch(gpu.ctx, n, s.as_ptr(), p.as_ptr(), out.as_mut_ptr(), st.as_mut_ptr(), 0) })?;
                // an undecodable input leaves 32 zero bytes, which would decode as a valid order-4 point: report it instead
                if let Some(i) = st.iter().position(|&x| x != 0) {
                    return Err(GpuError::InvalidPoint(i));
                }
            }
        }
        Ok(unpack_points(&out))
    }
}

fn status_to_result(st: u8) -> Result<(), SignatureError> {
    match st {
        0 => Ok(()),
        1 => Err(SignatureError::InvalidSignatureLength("expect 64".to_owned())),
        2 => Err(SignatureError::SignatureNotCanonical),
        3 => Err(SignatureError::RNotCanonical),
        4 => Err(SignatureError::RSmallOrder),
        5 => Err(SignatureError::PublicKeyNotCanonical),
        6 => Err(SignatureError::PublicKeySmallOrder),
        7 => Err(SignatureError::MarshallingError(MarshallingError::InvalidInput("invalid Ed25519 curve point".to_owned()))),
        _ => Err(SignatureError::InvalidSignature("reconstructed S is not equal to signature".to_owned())),
    }
}

/// Batch companion of `eddsa::verify_with_checks` (sign/eddsa/eddsa_sig.rs:159-212) and, with
/// `schnorr = true`, of `schnorr::verify_with_checks` (sign/schnorr/schnorr_sig.rs:53-110).
/// Items are (public key bytes, message, signature).
pub fn verify_batch(gpu: &Gpu, items: &[(&[u8], &[u8], &[u8])], schnorr: bool) -> Result<Vec<Result<(), SignatureError>>, GpuError> {
    let mut res: Vec<Option<Result<(), SignatureError>>> = vec![None; items.len()];
    let (mut pk, mut sig, mut msg, mut off, mut idx) = (vec![], vec![], vec![], vec![0u64], vec![]);
    for (i, (p, m, s)) in items.iter().enumerate() {
        if s.len() != 64 {
            res[i] = Some(status_to_result(1)); // eddsa_sig.rs:161 / schnorr_sig.rs:68
            continue;
        }
        if p.len() != 32 {
            // the reference cannot even build such a key: unmarshal_binary fails; one wrong-length key must not
            // misalign every later item of the packed batch
            res[i] = Some(status_to_result(5));
            continue;
        }
        pk.extend_from_slice(p);
        sig.extend_from_slice(s);
        msg.extend_from_slice(m);
        off.push(msg.len() as u64);
        idx.push(i);
    }
    let n = idx.len();
    let mut st = vec![0u8; n];
    let f = if schnorr { sys::kb_schnorr_verify_batch } else { sys::kb_eddsa_verify_batch };
    gpu.check(unsafe { f(gpu.ctx, n, pk.as_ptr(), msg.as_ptr(), off.as_ptr(), sig.as_ptr(), st.as_mut_ptr()) })?;
    for (k, i) in idx.into_iter().enumerate() {
        res[i] = Some(status_to_result(st[k]));
    }
    Ok(res.into_iter().map(|r| r.unwrap()).collect())
}

/// Batch companion of `PubPoly::eval` / `PubPoly::check` (share/poly.rs:457-469, :526-530) — the
/// group math of `Aggregator::verify_deal` (share/vss/pedersen/vss.rs:899-912).
pub trait PubPolyBatch {
    fn eval_batch(&self, gpu: &Gpu, idx: &[u32]) -> Result<Vec<EdPoint>, GpuError>;
    fn check_batch(&self, gpu: &Gpu, shares: &[PriShare<EdScalar>]) -> Result<Vec<bool>, GpuError>;
}
impl<G: kyber_rs::Group<POINT = EdPoint>> PubPolyBatch for PubPoly<G> {
    fn eval_batch(&self, gpu: &Gpu, idx: &[u32]) -> Result<Vec<EdPoint>, GpuError> {
        let (_, commits) = self.info();
        let c = pack_points(&commits);
        let pid = vec![0u32; idx.len()];
        let mut out = vec![0u8; 32 * idx.len()];
        let mut st = vec![0u8; idx.len()];
        gpu.check(unsafe { sys::kb_pubpoly_eval_batch(gpu.ctx, 1, commits.len(), c.as_ptr(), idx.len(), pid.as_ptr(), idx.as_ptr(), out.as_mut_ptr(), st.as_mut_ptr()) })?;
        if let Some(i) = st.iter().position(|&x| x != 0) {
            return Err(GpuError::InvalidPoint(i));
        }
        Ok(unpack_points(&out))
    }
    fn check_batch(&self, gpu: &Gpu, shares: &[PriShare<EdScalar>]) -> Result<Vec<bool>, GpuError> {
        let (_, commits) = self.info();
        let c = pack_points(&commits);
        let idx: Vec<u32> = shares.iter().map(|s| s.i as u32).collect();
        let pid = vec![0u32; idx.len()];
        let sh: Vec<u8> = shares.iter().flat_map(|s| s.v.v).collect();
        let mut verdict = vec![0u8; idx.len()];
        gpu.check(unsafe { sys::kb_vss_verify_deals_batch(gpu.ctx, 1, commits.len(), c.as_ptr(), idx.len(), pid.as_ptr(), idx.as_ptr(), sh.as_ptr(), verdict.as_mut_ptr()) })?;
        Ok(verdict.into_iter().map(|v| v == 1).collect())
    }
}

/// `sum_i s_i * P_i` — what `PriPoly::commit` / `recover_commit` (share/poly.rs:195, :566) fold by hand.
pub fn msm(gpu: &Gpu, scalars: &[EdScalar], points: &[EdPoint]) -> Result<EdPoint, GpuError> {
    assert_eq!(scalars.len(), points.len());
    let (s, p) = (pack_scalars(scalars), pack_points(points));
    let mut out = [0u8; 32];
    let mut bad = 0u64;
    gpu.check(unsafe { sys::kb_msm(gpu.ctx, scalars.len(), s.as_ptr(), p.as_ptr(), out.as_mut_ptr(), std::ptr::null_mut(), &mut bad) })?;
    if bad != 0 {
        return Err(GpuError::InvalidPoint(0)); // the sum is undefined when an input does not decode
    }
    Ok(unpack_points(&out).remove(0))
}


/// One DKG deal-verification round for this node's view of `dealers` (share/dkg/pedersen/dkg.rs:513-597): every dealer's
/// committed polynomial against the share it sent to every verifier, the commitments handed over in their in-memory form
/// (no per-commitment `marshal_binary`).  `verdict[d * n + i]` is what `Aggregator::verify_deal` (vss.rs:899-912) decides.
pub fn dkg_verify_round(gpu: &Gpu, n: usize, polys: &[Vec<EdPoint>], shares: &[EdScalar]) -> Result<Vec<bool>, GpuError> {
    let t = polys.first().map(|p| p.len()).unwrap_or(0);
    assert!(polys.iter().all(|p| p.len() == t) && shares.len() == polys.len() * n);
    let limbs: Vec<i32> = polys.iter().flat_map(|p| pack_point_limbs(p)).collect();
    let sh = pack_scalars(shares);
    let mut verdict = vec![0u8; polys.len() * n];
    gpu.check(unsafe { sys::kb_dkg_verify_round_limbs(gpu.ctx, n, t, 0, polys.len(), limbs.as_ptr(), sh.as_ptr(), verdict.as_mut_ptr()) })?;
    Ok(verdict.into_iter().map(|v| v == 1).collect())
}
