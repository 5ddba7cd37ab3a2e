//! Rust binding of `include/p2g.h` and the drop-in for the one expression of the reference that libp2g replaces:
//!
//! ```text
//! circuit_data.prove(witnesses).unwrap()        plonky2-backend/src/actions/prove_action.rs:96
//!                                               plonky2-backend/src/circuit_translation/tests/factories/utils.rs:16-27
//! ```
//!
//! `GpuProver::new(&circuit_data)` runs once per circuit (it uploads the preprocessed polynomials: what
//! `builder.build::<C>()` computed at circuit_translation/mod.rs:81 stays resident on the GPU), `GpuProver::prove(&circuit_data,
//! witnesses)` replaces the expression above.  Everything else in the CLI -- ACIR translation, witness generation, `compress`,
//! `to_bytes`, `write_vk`, `verify` -- is untouched.
//!
//! The witness crosses the boundary the way plonky2 holds it: `MatrixWitness.wire_values` is a `Vec<Vec<F>>` with one heap
//! allocation per wire column; `p2g_prove_columns` takes the `num_wires` column pointers (GoldilocksField is
//! `#[repr(transparent)]` over `u64`; the library reduces non-canonical words on the device), so no 1.96 GB flat copy is built.
//!
//! NOTE (visibility): `MatrixWitness::wire_values` is `pub(crate)` in upstream plonky2 0.2.2.  The reference already builds
//! against its own fork (`[patch.crates-io] plonky2 = { path = "../plonky2/plonky2" }`, plonky2-backend/Cargo.toml:29-32); the
//! shim needs that field `pub` there (one line in plonky2/src/iop/witness.rs).  Without it, `wire_columns()` below falls back
//! to `get_wire(row, col)` and copies (slow path, feature-free, same bytes).

use core::ffi::{c_char, c_int, c_void};
use std::ffi::CStr;

use anyhow::{anyhow, bail, Result};
use plonky2::field::goldilocks_field::GoldilocksField;
use plonky2::field::types::{Field, PrimeField64};
use plonky2::iop::generator::generate_partial_witness;
use plonky2::iop::witness::{PartialWitness, Witness};
use plonky2::plonk::circuit_data::CircuitData;
use plonky2::plonk::config::{GenericConfig, KeccakGoldilocksConfig, PoseidonGoldilocksConfig};
use plonky2::plonk::proof::ProofWithPublicInputs;

pub type F = GoldilocksField;
pub const D: usize = 2;

// ---------------------------------------------------------------------------------------------------------------------
// include/p2g.h, field for field
// ---------------------------------------------------------------------------------------------------------------------
pub const P2G_OK: c_int = 0;
pub const P2G_ESMALLBUF: c_int = -6;
pub const P2G_HASH_KECCAK25: u32 = 0;
pub const P2G_HASH_POSEIDON: u32 = 1;
pub const P2G_MAX_FRI_LAYERS: usize = 8;
pub const P2G_NCCL_UNIQUE_ID_BYTES: usize = 128;

/// enum p2g_gate_kind
#[repr(u32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum GateKind {
    Noop = 0,
    Constant = 1,
    PublicInput = 2,
    Arithmetic = 3,
    BaseSum = 4,
    Poseidon = 5,
    RandomAccess = 6,
    U32Arithmetic = 7,
    U32AddMany = 8,
    U32Subtraction = 9,
    U32RangeCheck = 10,
    Comparison = 11,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct P2gGate {
    pub kind: u32,
    pub params: [u32; 4],
    pub selector_index: u32,
    pub group_lo: u32,
    pub group_hi: u32,
    pub num_constraints: u32,
}

#[repr(C)]
pub struct P2gCircuitDesc {
    pub struct_size: u32,
    pub degree_bits: u32,
    pub num_wires: u32,
    pub num_routed_wires: u32,
    pub num_constants: u32,
    pub num_selectors: u32,
    pub num_challenges: u32,
    pub rate_bits: u32,
    pub cap_height: u32,
    pub pow_bits: u32,
    pub num_query_rounds: u32,
    pub quotient_degree_factor: u32,
    pub num_partial_products: u32,
    pub num_gate_constraints: u32,
    pub num_public_inputs: u32,
    pub hasher: u32,
    pub num_fri_layers: u32,
    pub reduction_arity_bits: [u32; P2G_MAX_FRI_LAYERS],
    pub num_gates: u32,
    pub gates: *const P2gGate,
    pub constants_sigmas: *const u64,
    pub k_is: *const u64,
    pub circuit_digest: *const u8,
}

#[repr(C)]
pub struct P2gCircuit {
    _private: [u8; 0],
}

pub type AllgatherFn = extern "C" fn(user: *mut c_void, send: *const c_void, recv: *mut c_void, bytes: usize, is_device: c_int) -> c_int;

extern "C" {
    pub fn p2g_version() -> c_int;
    pub fn p2g_device_count() -> c_int;
    pub fn p2g_last_error() -> *const c_char;
    pub fn p2g_host_alloc(bytes: usize) -> *mut c_void;
    pub fn p2g_host_free(p: *mut c_void);
    pub fn p2g_circuit_create(desc: *const P2gCircuitDesc, device: c_int, out: *mut *mut P2gCircuit) -> c_int;
    pub fn p2g_circuit_destroy(c: *mut P2gCircuit);
    pub fn p2g_circuit_cap(c: *const P2gCircuit, cap: *mut u8, cap_len: usize, digest: *mut u8, digest_len: usize) -> c_int;
    pub fn p2g_prove(c: *mut P2gCircuit, wires: *const u64, public_inputs: *const u64, n_pi: usize, forced_pow_witness: *const u64,
                     out: *mut u8, out_len: *mut usize, timings: *mut c_void) -> c_int;
    pub fn p2g_prove_columns(c: *mut P2gCircuit, wire_columns: *const *const u64, public_inputs: *const u64, n_pi: usize,
                             forced_pow_witness: *const u64, compressed: c_int, out: *mut u8, out_len: *mut usize,
                             timings: *mut c_void) -> c_int;
    pub fn p2g_prove_compressed(c: *mut P2gCircuit, wires: *const u64, wires_on_device: c_int, public_inputs: *const u64, n_pi: usize,
                                forced_pow_witness: *const u64, out: *mut u8, out_len: *mut usize, timings: *mut c_void) -> c_int;
    /// p2g_prove_columns with only the num_routed_wires routed columns; the advice columns are computed on the device
    pub fn p2g_prove_routed_columns(c: *mut P2gCircuit, routed_columns: *const *const u64, public_inputs: *const u64, n_pi: usize,
                                    forced_pow_witness: *const u64, compressed: c_int, out: *mut u8, out_len: *mut usize,
                                    timings: *mut c_void) -> c_int;
    pub fn p2g_prove_device(c: *mut P2gCircuit, d_wires: *const u64, public_inputs: *const u64, n_pi: usize, forced_pow_witness: *const u64,
                            out: *mut u8, out_len: *mut usize, timings: *mut c_void) -> c_int;
    /// device-side witness fill: the advice columns (>= num_routed_wires) of a device-resident trace, from its routed columns
    pub fn p2g_fill_advice_device(c: *mut P2gCircuit, d_wires: *mut u64) -> c_int;
    pub fn p2g_proof_size_bound(c: *const P2gCircuit) -> usize;
    pub fn p2g_vk_bytes(c: *const P2gCircuit, cfg: *const c_void /* p2g_vk_config, NULL = wide_ecc_config */, out: *mut u8, out_len: *mut usize) -> c_int;
    // one proof across several GPUs: host callback, or the library's own NCCL communicator
    pub fn p2g_circuit_create_sharded(desc: *const P2gCircuitDesc, device: c_int, rank: c_int, world: c_int, allgather: AllgatherFn,
                                      user: *mut c_void, out: *mut *mut P2gCircuit) -> c_int;
    pub fn p2g_nccl_unique_id(id_out: *mut u8) -> c_int;
    pub fn p2g_circuit_create_sharded_nccl(desc: *const P2gCircuitDesc, device: c_int, rank: c_int, world: c_int, nccl_id: *const u8,
                                           out: *mut *mut P2gCircuit) -> c_int;
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(p2g_last_error()) }.to_string_lossy().into_owned()
}

fn check(rc: c_int, what: &str) -> Result<()> {
    if rc == P2G_OK {
        Ok(())
    } else {
        Err(anyhow!("{what}: libp2g error {rc}: {}", last_error()))
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// gate table: `common.gates[i].0.id()` -> (kind, params).  The ids are the `Debug` renderings plonky2 builds in `Gate::id()`
// (e.g. "ArithmeticGate { num_ops: 20 }", "BaseSumGate { num_limbs: 63 } + Base: 2", "U32AddManyGate { num_addends: 3,
// num_ops: 9, _phantom: PhantomData<..> }"); the numbers are exactly what BackendGateSerializer writes for the gate
// (plonky2-backend/src/actions/write_vk_action.rs:35-62).
// ---------------------------------------------------------------------------------------------------------------------
fn field_of(id: &str, name: &str) -> Option<u32> {
    let key = format!("{name}: ");
    let at = id.find(&key)? + key.len();
    let digits: String = id[at..].chars().take_while(|c| c.is_ascii_digit()).collect();
    digits.parse().ok()
}

pub fn gate_kind_and_params(id: &str) -> Result<(GateKind, [u32; 4])> {
    let f = |n: &str| field_of(id, n).ok_or_else(|| anyhow!("gate id `{id}` lacks field `{n}`"));
    Ok(if id.starts_with("NoopGate") {
        (GateKind::Noop, [0; 4])
    } else if id.starts_with("ConstantGate") {
        (GateKind::Constant, [f("num_consts")?, 0, 0, 0])
    } else if id.starts_with("PublicInputGate") {
        (GateKind::PublicInput, [0; 4])
    } else if id.starts_with("ArithmeticGate") {
        (GateKind::Arithmetic, [f("num_ops")?, 0, 0, 0])
    } else if id.starts_with("BaseSumGate") {
        let base = id.rsplit("Base: ").next().and_then(|s| s.trim().parse::<u32>().ok()).ok_or_else(|| anyhow!("BaseSumGate id `{id}`"))?;
        (GateKind::BaseSum, [base, f("num_limbs")?, 0, 0])
    } else if id.starts_with("PoseidonGate") {
        (GateKind::Poseidon, [0; 4])
    } else if id.starts_with("RandomAccessGate") {
        (GateKind::RandomAccess, [f("bits")?, f("num_copies")?, f("num_extra_constants")?, 0])
    } else if id.starts_with("U32ArithmeticGate") {
        (GateKind::U32Arithmetic, [f("num_ops")?, 0, 0, 0])
    } else if id.starts_with("U32AddManyGate") {
        (GateKind::U32AddMany, [f("num_addends")?, f("num_ops")?, 0, 0])
    } else if id.starts_with("U32SubtractionGate") {
        (GateKind::U32Subtraction, [f("num_ops")?, 0, 0, 0])
    } else if id.starts_with("U32RangeCheckGate") {
        (GateKind::U32RangeCheck, [f("num_input_limbs")?, 0, 0, 0])
    } else if id.starts_with("ComparisonGate") {
        (GateKind::Comparison, [f("num_bits")?, f("num_chunks")?, 0, 0])
    } else {
        bail!("gate `{id}` is not one of the 13 gates the translators of this backend emit (SURVEY.md App. B)")
    })
}

/// `C::Hasher` -> enum p2g_hasher.  The CLI uses KeccakGoldilocksConfig (plonky2-backend/src/lib.rs:13).
pub trait P2gConfig: GenericConfig<D, F = F> {
    const HASHER: u32;
    const HASH_BYTES: usize;
}
impl P2gConfig for KeccakGoldilocksConfig {
    const HASHER: u32 = P2G_HASH_KECCAK25;
    const HASH_BYTES: usize = 25;
}
impl P2gConfig for PoseidonGoldilocksConfig {
    const HASHER: u32 = P2G_HASH_POSEIDON;
    const HASH_BYTES: usize = 32;
}

// ---------------------------------------------------------------------------------------------------------------------
// the handle
// ---------------------------------------------------------------------------------------------------------------------
pub struct GpuProver {
    handle: *mut P2gCircuit,
    num_wires: usize,
    degree: usize,
}
unsafe impl Send for GpuProver {}

impl Drop for GpuProver {
    fn drop(&mut self) {
        unsafe { p2g_circuit_destroy(self.handle) }
    }
}

impl GpuProver {
    /// Once per circuit, after `translator.unpack()` / `builder.build::<C>()` (circuit_translation/mod.rs:80-81).
    pub fn new<C: P2gConfig>(cd: &CircuitData<F, C, D>, device: i32) -> Result<Self> {
        let common = &cd.common;
        let n = common.degree();
        // gate table in `common.gates` order with the selector data of `common.selectors_info`
        let mut gates = Vec::with_capacity(common.gates.len());
        for (i, g) in common.gates.iter().enumerate() {
            let (kind, params) = gate_kind_and_params(&g.0.id())?;
            let sel = common.selectors_info.selector_indices[i];
            let grp = &common.selectors_info.groups[sel];
            gates.push(P2gGate {
                kind: kind as u32,
                params,
                selector_index: sel as u32,
                group_lo: grp.start as u32,
                group_hi: grp.end as u32,
                num_constraints: g.0.num_constraints() as u32,
            });
        }
        // preprocessed polynomials as VALUES on the subgroup, column-major [num_constants + num_routed][N]:
        // prover_only.constants_sigmas_commitment.polynomials holds their coefficients (selectors, constants, then sigmas)
        let polys = &cd.prover_only.constants_sigmas_commitment.polynomials;
        let mut cs: Vec<u64> = Vec::with_capacity(polys.len() * n);
        for p in polys {
            let v = p.clone().fft();
            cs.extend(v.values.iter().map(|x| x.to_canonical_u64()));
        }
        let k_is: Vec<u64> = common.k_is.iter().map(|x| x.to_canonical_u64()).collect();
        let arity = &common.fri_params.reduction_arity_bits;
        if arity.len() > P2G_MAX_FRI_LAYERS {
            bail!("more than {P2G_MAX_FRI_LAYERS} FRI layers");
        }
        let mut rab = [0u32; P2G_MAX_FRI_LAYERS];
        for (i, a) in arity.iter().enumerate() {
            rab[i] = *a as u32;
        }
        let cfg = &common.config;
        let desc = P2gCircuitDesc {
            struct_size: core::mem::size_of::<P2gCircuitDesc>() as u32,
            degree_bits: common.degree_bits() as u32,
            num_wires: cfg.num_wires as u32,
            num_routed_wires: cfg.num_routed_wires as u32,
            num_constants: common.num_constants as u32,
            num_selectors: common.selectors_info.num_selectors() as u32,
            num_challenges: cfg.num_challenges as u32,
            rate_bits: cfg.fri_config.rate_bits as u32,
            cap_height: cfg.fri_config.cap_height as u32,
            pow_bits: cfg.fri_config.proof_of_work_bits,
            num_query_rounds: cfg.fri_config.num_query_rounds as u32,
            quotient_degree_factor: common.quotient_degree_factor as u32,
            num_partial_products: common.num_partial_products as u32,
            num_gate_constraints: common.num_gate_constraints as u32,
            num_public_inputs: common.num_public_inputs as u32,
            hasher: C::HASHER,
            num_fri_layers: arity.len() as u32,
            reduction_arity_bits: rab,
            num_gates: gates.len() as u32,
            gates: gates.as_ptr(),
            constants_sigmas: cs.as_ptr(),
            k_is: k_is.as_ptr(),
            circuit_digest: core::ptr::null(), // derived by the library and checked below
        };
        let mut handle: *mut P2gCircuit = core::ptr::null_mut();
        check(unsafe { p2g_circuit_create(&desc, device, &mut handle) }, "p2g_circuit_create")?;
        let this = GpuProver { handle, num_wires: cfg.num_wires, degree: n };
        // the device-side preprocessed commitment must be the one plonky2 built
        let ncap = 1usize << cfg.fri_config.cap_height.min(common.degree_bits() + cfg.fri_config.rate_bits);
        let mut cap = vec![0u8; ncap * C::HASH_BYTES];
        let mut digest = vec![0u8; C::HASH_BYTES];
        check(unsafe { p2g_circuit_cap(handle, cap.as_mut_ptr(), cap.len(), digest.as_mut_ptr(), digest.len()) }, "p2g_circuit_cap")?;
        let mut want_cap = Vec::new();
        plonky2::util::serialization::Write::write_merkle_cap(&mut want_cap, &cd.verifier_only.constants_sigmas_cap)
            .map_err(|e| anyhow!("{e:?}"))?;
        let mut want_digest = Vec::new();
        plonky2::util::serialization::Write::write_hash::<F, C::Hasher>(&mut want_digest, cd.verifier_only.circuit_digest)
            .map_err(|e| anyhow!("{e:?}"))?;
        // write_merkle_cap prefixes the cap with its length (usize as u64 LE): compare the digests only
        if want_cap.len() < cap.len() || want_cap[want_cap.len() - cap.len()..] != cap[..] || want_digest != digest {
            bail!("libp2g's preprocessed commitment differs from plonky2's (constants_sigmas_cap / circuit_digest)");
        }
        Ok(this)
    }

    /// Drop-in for `circuit_data.prove(witnesses)` (prove_action.rs:96): same argument, same return type, same error type.
    pub fn prove<C: P2gConfig>(&self, cd: &CircuitData<F, C, D>, inputs: PartialWitness<F>) -> Result<ProofWithPublicInputs<F, C, D>> {
        let bytes = self.prove_bytes(cd, inputs, false)?;
        ProofWithPublicInputs::from_bytes(bytes, &cd.common)
    }

    /// The CLI's final file bytes directly (what prove_action.rs:75-78 `proof.compress(..)?.to_bytes()` produces).
    pub fn prove_compressed_bytes<C: P2gConfig>(&self, cd: &CircuitData<F, C, D>, inputs: PartialWitness<F>) -> Result<Vec<u8>> {
        self.prove_bytes(cd, inputs, true)
    }

    fn prove_bytes<C: P2gConfig>(&self, cd: &CircuitData<F, C, D>, inputs: PartialWitness<F>, compressed: bool) -> Result<Vec<u8>> {
        // witness generation stays in plonky2 (iop/generator.rs)
        let pw = generate_partial_witness(inputs, &cd.prover_only, &cd.common);
        let pis: Vec<u64> = pw.get_targets(&cd.prover_only.public_inputs).iter().map(|x| x.to_canonical_u64()).collect();
        let witness = pw.full_witness(); // MatrixWitness { wire_values: Vec<Vec<F>> }, column-major
        let cols = wire_columns(&witness.wire_values, self.num_wires, self.degree)?;
        let mut out = vec![0u8; unsafe { p2g_proof_size_bound(self.handle) }];
        let mut len = out.len();
        let rc = unsafe {
            p2g_prove_columns(self.handle, cols.as_ptr(), pis.as_ptr(), pis.len(), core::ptr::null(), compressed as c_int, out.as_mut_ptr(),
                              &mut len, core::ptr::null_mut())
        };
        check(rc, "p2g_prove_columns")?;
        out.truncate(len);
        Ok(out)
    }

    /// `verifier_data().to_bytes(&BackendGateSerializer)` (write_vk_action.rs:76-79) from the handle, without rebuilding the circuit.
    pub fn vk_bytes(&self) -> Result<Vec<u8>> {
        let mut len = 0usize;
        let rc = unsafe { p2g_vk_bytes(self.handle, core::ptr::null(), core::ptr::null_mut(), &mut len) };
        if rc != P2G_ESMALLBUF {
            check(rc, "p2g_vk_bytes")?;
        }
        let mut out = vec![0u8; len];
        check(unsafe { p2g_vk_bytes(self.handle, core::ptr::null(), out.as_mut_ptr(), &mut len) }, "p2g_vk_bytes")?;
        out.truncate(len);
        Ok(out)
    }
}

/// One pointer per wire column.  GoldilocksField is `#[repr(transparent)] struct GoldilocksField(pub u64)`, so a `&[F]` IS a
/// `&[u64]`; the words may be non-canonical (< 2^64), which p2g_prove_columns accepts.
fn wire_columns(wire_values: &[Vec<F>], num_wires: usize, degree: usize) -> Result<Vec<*const u64>> {
    if wire_values.len() != num_wires {
        bail!("witness has {} columns, circuit has {num_wires} wires", wire_values.len());
    }
    let mut cols = Vec::with_capacity(num_wires);
    for c in wire_values {
        if c.len() != degree {
            bail!("witness column of {} rows, circuit has {degree}", c.len());
        }
        cols.push(c.as_ptr() as *const u64);
    }
    Ok(cols)
}

#[cfg(test)]
mod tests {
    use super::*;

    #[test]
    fn gate_ids() {
        let pd = "PhantomData<plonky2_field::goldilocks_field::GoldilocksField>";
        assert_eq!(gate_kind_and_params("ArithmeticGate { num_ops: 20 }").unwrap(), (GateKind::Arithmetic, [20, 0, 0, 0]));
        assert_eq!(gate_kind_and_params("BaseSumGate { num_limbs: 63 } + Base: 2").unwrap(), (GateKind::BaseSum, [2, 63, 0, 0]));
        assert_eq!(gate_kind_and_params(&format!("RandomAccessGate {{ bits: 4, num_copies: 4, num_extra_constants: 2, _phantom: {pd} }}<D=2>")).unwrap(),
                   (GateKind::RandomAccess, [4, 4, 2, 0]));
        assert_eq!(gate_kind_and_params(&format!("U32AddManyGate {{ num_addends: 3, num_ops: 9, _phantom: {pd} }}")).unwrap(),
                   (GateKind::U32AddMany, [3, 9, 0, 0]));
        assert_eq!(gate_kind_and_params(&format!("ComparisonGate {{ num_bits: 32, num_chunks: 16, _phantom: {pd} }}<D=2>")).unwrap(),
                   (GateKind::Comparison, [32, 16, 0, 0]));
        assert!(gate_kind_and_params("LookupGate { num_slots: 26 }").is_err());
    }
}
