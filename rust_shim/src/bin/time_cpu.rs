//! `cargo run --release --features time-cpu --bin time_cpu -- <rows_log2> <repeats>`
//!
//! Times plonky2's OWN prover -- `prove_with_partition_witness`, Rayon on every host core, the function behind
//! `circuit_data.prove(..)` at plonky2-backend/src/actions/prove_action.rs:96 -- on an AssertZero-chain circuit of 2^rows_log2
//! rows built with the reference's configuration (`CircuitConfig::wide_ecc_config()`, circuit_translation/mod.rs:69;
//! `KeccakGoldilocksConfig`, lib.rs:13), and the same call through libp2g.  Prints one JSON line per arm in bench.py's format
//! (`"impl": "reference"`, `cpu_baseline.kind = "reference"`).  This is the CPU arm bench.py cannot run in the development image
//! (no Rust toolchain there); with it `oracle/_ref` semantics are met by the real thing.
use std::time::Instant;

use plonky2::field::types::Field;
use plonky2::iop::generator::generate_partial_witness;
use plonky2::iop::witness::{PartialWitness, WitnessWrite};
use plonky2::plonk::circuit_builder::CircuitBuilder;
use plonky2::plonk::circuit_data::CircuitConfig;
use plonky2::plonk::config::KeccakGoldilocksConfig;
use plonky2::plonk::prover::prove_with_partition_witness;
use plonky2::util::timing::TimingTree;

use p2g_shim::{GpuProver, D, F};

type C = KeccakGoldilocksConfig;

fn main() -> anyhow::Result<()> {
    let args: Vec<String> = std::env::args().collect();
    let rows_log2: usize = args.get(1).map(|s| s.parse().unwrap()).unwrap_or(16);
    let repeats: usize = args.get(2).map(|s| s.parse().unwrap()).unwrap_or(3);
    // AssertZero chain: x_{i+1} = x_i * x_i + x_i, 20 ArithmeticGate ops per row (BASELINE.json configs[1])
    let mut builder = CircuitBuilder::<F, D>::new(CircuitConfig::wide_ecc_config());
    let x0 = builder.add_virtual_target();
    let mut x = x0;
    let ops = (1usize << rows_log2) * 20 * 97 / 100;
    for _ in 0..ops {
        x = builder.mul_add(x, x, x);
    }
    builder.register_public_input(x);
    let cd = builder.build::<C>();
    assert_eq!(cd.common.degree_bits(), rows_log2, "adjust the op count: the circuit did not land on 2^{rows_log2} rows");
    let cores = std::thread::available_parallelism().map(|n| n.get()).unwrap_or(1);

    let mut cpu_s = Vec::new();
    for _ in 0..repeats {
        let mut pw = PartialWitness::<F>::new();
        pw.set_target(x0, F::from_canonical_u64(3));
        let partition = generate_partial_witness(pw, &cd.prover_only, &cd.common); // witness generation outside the timed call
        let mut timing = TimingTree::default();
        let t = Instant::now();
        let proof = prove_with_partition_witness(&cd.prover_only, &cd.common, partition, &mut timing)?;
        cpu_s.push(t.elapsed().as_secs_f64());
        cd.verify(proof)?;
    }
    let best = cpu_s.iter().cloned().fold(f64::INFINITY, f64::min);
    println!(
        "{{\"impl\": \"reference\", \"metric\": \"proofs/sec\", \"value\": {:.6}, \"unit\": \"proofs/s\", \"ms_per_step\": {:.3}, \
         \"config\": {{\"workload\": \"assert_zero_2^{rows_log2}\", \"rows\": {}, \"wires\": 234, \"hasher\": \"keccak25\"}}, \
         \"cpu_baseline\": {{\"kind\": \"reference\", \"cores\": {cores}, \"sample\": \"plonky2 prove_with_partition_witness, Rayon, best of {repeats}\"}}}}",
        1.0 / best, best * 1e3, 1usize << rows_log2
    );

    let gpu = GpuProver::new(&cd, 0)?;
    let mut gpu_s = Vec::new();
    for _ in 0..repeats {
        let mut pw = PartialWitness::<F>::new();
        pw.set_target(x0, F::from_canonical_u64(3));
        let t = Instant::now();
        let proof = gpu.prove(&cd, pw)?; // includes witness generation, like circuit_data.prove
        gpu_s.push(t.elapsed().as_secs_f64());
        cd.verify(proof)?; // plonky2's verifier accepts libp2g's proof
    }
    let best = gpu_s.iter().cloned().fold(f64::INFINITY, f64::min);
    println!("{{\"impl\": \"p2g\", \"value\": {:.6}, \"unit\": \"proofs/s\", \"ms_per_step\": {:.3}, \"note\": \"includes generate_partial_witness\"}}",
             1.0 / best, best * 1e3);
    Ok(())
}
