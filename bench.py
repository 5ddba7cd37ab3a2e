#!/usr/bin/env python
"""bench.py -- proofs/sec of the B200 prover on the BASELINE workload (2^20-row ECDSA-shaped ACIR circuit, 234 wires,
KeccakGoldilocksConfig), with the NTT / Merkle rooflines and the CPU prover timed beside it.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the CPU restatement of the reference prover (oracle port)

A "step" is one proof: witness matrix -> proof bytes.  `value` is measured with the witness resident in HBM
(p2g_prove_device), `e2e` through CircuitData.prove with the witness in pinned host memory (H2D inside the timed region,
proof bytes written to host memory).  Multi-GPU (torchrun, one rank per GPU): every rank proves independent witnesses of
the same circuit (weak scaling, no data-path collective).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "proofs/sec on a 2^20-row ACIR circuit (ECDSA-shaped, 234 wires, Keccak config)"
UNIT = "proofs/s"
# constants read off the committed ncu --set full captures (profiles/r1_summary.md section 3)
NCU = {"lde_traffic_over_algorithmic": 6.26 / 2.265, "keccak_traffic_over_algorithmic": 15.98 / 15.70,
       "lde_limiter": "integer issue: 413 thread instructions per output element, issue slots 59-60% busy (ALU pipe 58-62%, FMA pipe 22-25%), DRAM 1.3 TB/s: "
                      "64-bit modular arithmetic on 32-bit pipes, 59% of the issue roofline, not HBM-bound",
       "keccak_limiter": "integer ALU pipe 99.7% active (LOP3/SHF): at the hardware floor for Keccak-f",
       "quotient_dram_bytes": 65.5e9,
       "quotient_limiter": "integer issue: 84 k instructions per point in the limb sweep (8.1 k IMAD.WIDE), issue slots 55 % busy, DRAM 0.6-2.8 TB/s",
       "files": ["profiles/r2_ntt_ncu_raw.csv", "profiles/r2_keccak_ncu_raw.csv", "profiles/r2_quotient_ncu_raw.csv", "profiles/r2_launches_bench.csv"]}


REAL_WORKLOAD = "ecdsa-real"   # the EcdsaSecp256k1 opcode translated like the reference does (acir/ecdsa.h), not a synthetic gate mix


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="p2g", choices=["p2g", "reference"])
    ap.add_argument("--degree-bits", type=int, default=20)
    ap.add_argument("--workload", default=REAL_WORKLOAD,
                    help="ecdsa-real (default): BASELINE configs[3] from real opcodes -- the Noir signature-check program on 2^(bits-17) "
                         "EcdsaSecp256k1 calls, translated like the reference does; ecdsa / sha256 / assert_zero / range: synthetic "
                         "gate mixes of those shapes (p2g.synth), any size")
    ap.add_argument("--hasher", default="keccak25")
    ap.add_argument("--inflight", type=int, default=2,
                    help="proofs in flight per GPU (one circuit handle + stream + host thread each); a step is still one proof")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable-columns e2e measurement")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the one-proof-across-N-GPUs measurement")
    ap.add_argument("--sharded-callback", action="store_true",
                    help="N > 1: exchange through the host callback bound to torch.distributed instead of the library's own NCCL communicator")
    ap.add_argument("--cpu-sample-bits", type=int, default=0, help="rows (log2) of the CPU-baseline sample; 0 = auto")
    ap.add_argument("--no-full-size-cpu", action="store_true", help="--impl reference: skip the one full-size CPU proof (samples only)")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_workload(p2g, args, bits, seed, pinned=False):
    """The circuit + witness of `--workload` at 2^bits rows: a synthetic gate mix (p2g.synth) or, for `ecdsa-real`, the Noir
    signature-check program on 2^(bits - 17) signatures translated from real opcodes."""
    cfg = p2g.CircuitConfig.wide_ecc_config(hasher=args.hasher)
    if args.workload == REAL_WORKLOAD:
        return p2g.ecdsa_inputs.RealEcdsaCircuit(bits, p2g.acir, config=cfg, seed=seed, pinned=pinned)
    return p2g.synth.SyntheticCircuit(bits, args.workload, config=cfg, num_public_inputs=4, seed=seed, pinned=pinned)


def data_note(args):
    return "synthetic (deterministic keys and signatures; circuit from the real opcode)" if args.workload == REAL_WORKLOAD else "synthetic"


def sample_note(args, bits):
    """What a CPU sample at 2^bits rows is, in words."""
    if args.workload == REAL_WORKLOAD:
        return f"the same program on {1 << (bits - 17)} signature(s), 2^{bits} rows"
    return f"the same gate mix at 2^{bits} rows"


def min_sample_bits(args):
    return 17 if args.workload == REAL_WORKLOAD else 11


def workload_config(args, sc):
    """The `config` object of the JSON line: identical for the product arm and the reference arm (same workload, same circuit)."""
    cfg = sc.config
    return {"workload": f"{args.workload}_2^{args.degree_bits}", "rows": 1 << args.degree_bits, "wires": cfg.num_wires,
            "routed": cfg.num_routed_wires, "hasher": args.hasher, "rate_bits": cfg.rate_bits, "gates": len(sc.common.gates),
            "gate_constraints": sc.common.num_gate_constraints, "fri_arity_bits": list(sc.common.reduction_arity_bits),
            "public_inputs": len(sc.public_inputs),
            **({"source": f"{sc.num_signatures} EcdsaSecp256k1 opcode(s) + 160 RANGE opcodes each, translated like "
                          "circuit_translation/ecdsa_secp256k1_translator.rs; witness from the restated plonky2_ecdsa generators",
                "rows_used": sc.rows_used} if args.workload == REAL_WORKLOAD else {}),
            "l2": "inputs larger than L2 (1.96 GB trace, 14.6 GB LDE per proof at 2^20 rows); no explicit flush"}


def cpu_proof(p2g, args, bits, seed, sc=None):
    """One oracle (CPU port) proof of the workload's gate mix at 2^bits rows.  Returns (seconds, cores, circuit)."""
    from oracle import corc
    corc.lib(native=True)      # -march=native build, compiled on this host (oracle/corc.py build_native)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_cd
    if sc is None:
        sc = make_workload(p2g, args, bits, seed)
    cd = oracle_cd(sc.common)
    op = corc.OracleProver(cd, sc.constants_sigmas)     # circuit build (preprocessed commitment) is outside the timing
    t0 = time.perf_counter()
    op.prove(sc.wires, sc.public_inputs)
    dt = time.perf_counter() - t0
    op.close()
    return dt, corc.num_threads(), sc


def pick_cpu_sample_bits(p2g, args, target_s):
    if args.cpu_sample_bits:
        return max(args.cpu_sample_bits, min_sample_bits(args)) if args.workload == REAL_WORKLOAD else args.cpu_sample_bits
    bits = min(min_sample_bits(args), args.degree_bits)
    dt, _, _ = cpu_proof(p2g, args, bits, 1)
    # prover cost is ~linear in rows
    while bits < args.degree_bits and dt * 2 <= target_s:
        dt *= 2
        bits += 1
    return bits


def host_ram_gb():
    try:
        with open("/proc/meminfo") as f:
            for ln in f:
                if ln.startswith("MemAvailable:"):
                    return int(ln.split()[1]) / (1 << 20)
    except OSError:
        pass
    return 0.0


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (the oracle's C/OpenMP port of plonky2's prover, every host thread) on the
    arm's own config.  The FIRST timed step is one real proof of the full-size circuit (2^degree_bits rows): `value` is 1 / its
    measured time, not an extrapolation.  The remaining K-1 steps and the warm-up are bounded samples (the same gate mix at fewer
    rows) so that the run ends within minutes; their measured scaling factor against the full-size proof is printed beside it."""
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the CPU arm is meant to use every host thread
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from __graft_entry__ import load_product
    p2g = load_product()
    from oracle import corc
    corc.lib(native=True)
    bits = pick_cpu_sample_bits(p2g, args, 2.5)
    for i in range(args.warmup):
        cpu_proof(p2g, args, bits, 100 + i)
    cfg = p2g.CircuitConfig.wide_ecc_config(hasher=args.hasher)
    full_sc = make_workload(p2g, args, args.degree_bits, 0xAC1D + 3)
    need_gb = 45.0 * (1 << args.degree_bits) / (1 << 20)
    full_ok = host_ram_gb() >= need_gb and not args.no_full_size_cpu
    times, full_s, cores = [], None, 1
    for i in range(args.steps):
        if i == 0 and full_ok:
            dt, cores, _ = cpu_proof(p2g, args, args.degree_bits, 0, sc=full_sc)
            full_s = dt
        else:
            dt, cores, _ = cpu_proof(p2g, args, bits, 200 + i)
        times.append(dt)
    samples = times[1:] if full_s is not None else times
    sample_s = sum(samples) / len(samples) if samples else None
    nominal = float(1 << (args.degree_bits - bits))
    if full_s is not None:
        per_proof_s = full_s
        how = (f"step 1 = ONE REAL 2^{args.degree_bits}-row proof ({full_s:.1f} s: value = 1 / that, measured, same circuit as the "
               f"product arm); steps 2..{args.steps} and the warm-up = {sample_note(args, bits)} "
               f"({(sample_s or 0):.2f} s each)")
    else:
        per_proof_s = sample_s * nominal
        how = (f"host memory too small for the full-size oracle proof ({need_gb:.0f} GB needed): every step is a 2^{bits}-row "
               f"sample, time scaled x{int(nominal)} (extrapolated)")
    value = 1.0 / per_proof_s
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sum(times) / len(times) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 (Goldilocks field, 64-bit integer)", "data": data_note(args),
            "config": workload_config(args, full_sc),
            "full_size_ms": full_s * 1e3 if full_s is not None else None, "sample_rows_log2": bits,
            "sample_ms": sample_s * 1e3 if sample_s is not None else None,
            "measured_scaling_factor": (full_s / sample_s) if (full_s is not None and sample_s) else None,
            "nominal_scaling_factor": nominal,
            "ms_per_step_note": "mean wall time of the K timed steps as executed (one full-size proof + K-1 bounded samples); "
                                "`value` is proofs/s of the full-size proof alone",
            "reference_note": "the reference Rust prover (plonky2 0.2.2 fork + Rayon) cannot be built here: no cargo, crate not vendored "
                              "(DESIGN.md section 3); this is the C/OpenMP restatement of the same algorithm (oracle/c), " + corc.build_note(),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": how},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """The one JSON line, on the process's real stdout."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    args = parse()
    # stdout carries exactly one JSON line: everything libraries print (e.g. the NCCL version banner, written straight to fd 1 when
    # the communicator is created) is sent to stderr instead
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_product
    p2g = load_product()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = p2g.CircuitConfig.wide_ecc_config(hasher=args.hasher)
    if world > 1:   # input generation (untimed): this rank's share of the host cores, not torchrun's OMP_NUM_THREADS=1
        p2g.synth.set_threads(max(1, (os.cpu_count() or world) // world))
        p2g.acir.set_threads(max(1, (os.cpu_count() or world) // world))
    sc = make_workload(p2g, args, args.degree_bits, 0xAC1D + 3 + 1000 * rank, pinned=True)
    # circuit build: once, outside the timing.  `inflight` handles = proofs in flight on this GPU (own stream + host thread each)
    F = max(1, args.inflight)
    handles = [p2g.CircuitData(sc.common, sc.constants_sigmas, device=local_rank) for _ in range(F)]
    data = handles[0]
    wires_host = sc._wires_t                                  # pinned host tensor (int64 bit pattern of canonical u64)
    wires_dev = wires_host.cuda(non_blocking=False)
    h2d_bytes = wires_host.numel() * 8

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; device-side elapsed via CUDA events; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        outs = fn(steps)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), wall * 1e3)           # prove() is synchronous; both clocks cover the same region
        barrier()
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, outs

    def prove_steps(w, columns=False):
        """steps -> list of proofs: `steps` proofs in total, spread over the F handles, each driven by its own host thread.
        columns: w is a list of separately allocated pageable wire columns (p2g_prove_columns)."""
        def one(h):
            if columns == "routed":
                return h.prove_routed_columns(w, sc.public_inputs)
            return h.prove_columns(w, sc.public_inputs) if columns else h.prove(w, sc.public_inputs)

        def run(steps):
            if F == 1:
                return [one(data) for _ in range(steps)]
            outs, errs = [[] for _ in range(F)], []

            def work(i):
                try:
                    torch.cuda.set_device(local_rank)
                    for _ in range(steps // F + (1 if i < steps % F else 0)):
                        outs[i].append(one(handles[i]))
                except BaseException as e:  # noqa: BLE001
                    errs.append(e)
            ts = [threading.Thread(target=work, args=(i,)) for i in range(F)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            if errs:
                raise errs[0]
            return [o for lst in outs for o in lst]
        return run

    prove_steps(wires_dev)(args.warmup * F)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, outs = timed(prove_steps(wires_dev), args.steps)
    prove_steps(wires_host)(min(args.warmup, 2) * F)
    ms_e2e, outs_e2e = timed(prove_steps(wires_host), args.steps)
    # the same call from ordinary (pageable) memory laid out as plonky2 holds the witness -- one heap allocation per wire column
    # (MatrixWitness.wire_values), handed over as column pointers: what the Rust shim does, with no flat copy and no pinning
    ms_pageable = ms_routed = None
    if world == 1 and not args.no_pageable:
        cols = [sc.wires[i].copy() for i in range(sc.wires.shape[0])]
        prove_steps(cols, True)(F)
        ms_pageable, outs_pg = timed(prove_steps(cols, True), max(2, args.steps // 2))
        ms_pageable /= max(2, args.steps // 2)
        assert outs_pg[0].to_bytes() == outs_e2e[0].to_bytes()
        # ... and with only the routed columns handed over (p2g_prove_routed_columns): the advice columns, two thirds of the trace,
        # are computed on the device inside the upload pipeline.  An extra: a failure here must not cost the bench line.
        ms_routed = None
        try:
            rcols = cols[:cfg.num_routed_wires]
            prove_steps(rcols, "routed")(F)
            ms_routed, outs_rt = timed(prove_steps(rcols, "routed"), max(2, args.steps // 2))
            ms_routed /= max(2, args.steps // 2)
            if outs_rt[0].to_bytes() != outs_e2e[0].to_bytes():
                ms_routed = None
        except Exception as e:  # noqa: BLE001
            print(f"bench.py: routed-columns measurement skipped: {e}", file=sys.stderr)
            ms_routed = None
        del cols
    # per-kernel / per-stage device timings (CUDA events on the library's stream): with several proofs in flight the stage
    # events of one proof span kernels of the others, so they are taken from two proofs run alone right after the timed region
    solo = [data.prove(wires_dev, sc.public_inputs) for _ in range(2)] if F > 1 else outs
    clocks = sampler.stop() if rank == 0 else None

    # one proof across all N GPUs (coset sharding + NCCL all-gathers): latency of a single proof
    sharded = None
    if world > 1 and not args.no_sharded:
        for hd in handles:
            hd.close()
        grp = (p2g.sharding.TorchDistGroup(device=local_rank) if args.sharded_callback
               else p2g.sharding.NcclGroup.from_torch_dist(local_rank))
        sc0 = sc if rank == 0 else None
        if rank != 0:   # every rank proves the SAME circuit and witness
            sc0 = make_workload(p2g, args, args.degree_bits, 0xAC1D + 3, pinned=True)
        sdata = p2g.CircuitData(sc0.common, sc0.constants_sigmas, device=local_rank, shard=grp)
        shard_info = sdata.read(p2g.lib.BUF_SHARD_INFO)
        wh = sc0._wires_t
        wd = wh.cuda()
        for _ in range(args.warmup):
            sdata.prove(wd, sc0.public_inputs)
        ms_sd, souts = timed(lambda k: [sdata.prove(wd, sc0.public_inputs) for _ in range(k)], args.steps)
        sdata.prove(wh, sc0.public_inputs)
        ms_se, souts_e = timed(lambda k: [sdata.prove(wh, sc0.public_inputs) for _ in range(k)], args.steps)
        shard_info1 = sdata.read(p2g.lib.BUF_SHARD_INFO)
        sdata.prove(wd, sc0.public_inputs)
        shard_info2 = sdata.read(p2g.lib.BUF_SHARD_INFO)
        digest = torch.tensor(list(__import__("hashlib").sha256(souts[0].to_bytes()).digest()), device="cuda", dtype=torch.int32)
        gathered = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(gathered, digest)
        same = all(bool((g == gathered[0]).all()) for g in gathered)
        sharded = {"proofs_per_s": args.steps / (ms_sd / 1e3), "ms_per_proof": ms_sd / args.steps,
                   "e2e_ms_per_proof": ms_se / args.steps, "h2d_bytes_per_rank": int(souts_e[0].timings.get("h2d_bytes", 0)),
                   "collectives_per_proof": int(shard_info2[5] - shard_info1[5]) if len(shard_info1) > 5 else None,
                   "communicator": "library-owned NCCL (ncclAllGather on the handle's stream)" if int(shard_info[4]) else "host callback -> torch.distributed",
                   "identical_bytes_on_all_ranks": same,
                   "matches_single_gpu_bytes": (souts[0].to_bytes() == outs[0].to_bytes()) if rank == 0 else None,
                   "inverse_ntt_exchange": "peer stores over NVLink, fused into the transform" if int(shard_info[3]) else "NCCL all-gather",
                   "stages_ms": {k: round(sum(o.timings[k] for o in souts) / args.steps, 3)
                                 for k in ["wires_commit_ms", "zs_pp_ms", "quotient_ms", "openings_ms", "fri_ms", "total_ms"]}}
        sdata.close()

    if rank == 0:
        K = args.steps
        tms = [o.timings for o in solo]
        mean = lambda k: sum(t[k] for t in tms) / len(tms)
        hbm, src = peaks()
        leaf_gbs = sum(t["leaf_hash_bytes"] for t in tms) / 1e9 / (sum(t["leaf_hash_ms"] for t in tms) / 1e3)
        lde_gbs = sum(t["lde_bytes"] for t in tms) / 1e9 / (sum(t["lde_ms"] for t in tms) / 1e3)
        ntt_gbs = sum(t["ntt_bytes"] for t in tms) / 1e9 / (sum(t["ntt_ms"] for t in tms) / 1e3)
        stages = {k: round(mean(k), 3) for k in ["wires_commit_ms", "zs_pp_ms", "quotient_ms", "openings_ms", "fri_ms", "d2h_ms",
                                                 "total_ms", "ntt_ms", "merkle_ms", "leaf_hash_ms", "lde_ms",
                                                 "quotient_kernel_ms"]}
        q_alg = 8.0 * (sc.common.degree() << cfg.rate_bits) * (sc.common.num_preprocessed + cfg.num_wires + 20 + 2 + 2)
        launches = sum(o.timings["kernel_launches"] for o in outs) + sum(o.timings["kernel_launches"] for o in outs_e2e)
        proof_bytes = len(outs[0].to_bytes())
        line = {
            "metric": METRIC, "value": world * K / (ms_dev / 1e3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 (Goldilocks field, 64-bit integer)", "data": data_note(args),
            "config": workload_config(args, sc),
            "execution": {"parallelism": f"{world} GPU(s) x {F} proofs in flight per GPU (independent witnesses per GPU)",
                          "inflight_per_gpu": F, "proof_bytes": proof_bytes},
            "e2e": {"value": world * K / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / K, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": proof_bytes, "host_memory": "pinned (p2g_host_alloc), one [wires][rows] block",
                    "pageable_columns_ms_per_step": ms_pageable,
                    "pageable_routed_columns_ms_per_step": ms_routed,
                    "pageable_columns_note": "p2g_prove_columns from ordinary memory, one allocation per wire column (MatrixWitness.wire_values "
                                             "as the Rust shim passes it), staged through the library's pinned ring; measured at N = 1"},
            "gpu_launches": int(launches),
            # dominant stage of the step and the north-star's headline: the coset-LDE passes (NTT stage).  achieved = algorithmic
            # bytes (8N read + 64N written per column) / device time of those launches (CUDA events on the
            # library's stream); traffic = DRAM bytes actually moved, from the committed ncu --set full capture (2.73 x: two passes)
            "roofline": {"kernel": "k_pass_strided_fwd + k_pass_contig_fwd (rate-8 coset LDE of the trace / Z / quotient columns: the NTT stage)",
                         "bound": "hbm", "achieved": lde_gbs, "peak": hbm, "unit": "GB/s", "frac": lde_gbs / hbm,
                         "traffic": tms[0]["lde_bytes"] * NCU["lde_traffic_over_algorithmic"], "traffic_unit": "bytes per step",
                         "algorithmic_bytes_per_step": tms[0]["lde_bytes"], "peak_source": src,
                         "launches_per_step": tms[0]["lde_launches"],
                         "avg_launch_ms": sum(t["lde_ms"] for t in tms) / max(1, sum(t["lde_launches"] for t in tms)),
                         "share_of_step": sum(t["lde_ms"] for t in tms) / sum(t["total_ms"] for t in tms),
                         "limiter": NCU["lde_limiter"], "ncu": NCU["files"]},
            "ntt_roofline": {"kernel": "all NTT kernels (inverse NTTs + LDE passes)", "bound": "hbm", "achieved": ntt_gbs, "peak": hbm,
                             "unit": "GB/s", "frac": ntt_gbs / hbm, "algorithmic_bytes_per_step": tms[0]["ntt_bytes"]},
            "merkle_roofline": {"kernel": "k_leaf_keccak (Merkle leaf hashing of the LDE columns)", "bound": "hbm", "achieved": leaf_gbs,
                                "peak": hbm, "unit": "GB/s", "frac": leaf_gbs / hbm,
                                "traffic": tms[0]["leaf_hash_bytes"] * NCU["keccak_traffic_over_algorithmic"],
                                "traffic_unit": "bytes per step", "launches_per_step": tms[0]["leaf_hash_launches"],
                                "avg_launch_ms": sum(t["leaf_hash_ms"] for t in tms) / sum(t["leaf_hash_launches"] for t in tms),
                                "limiter": NCU["keccak_limiter"]},
            # constraint evaluation (quotient.cu): bounded by integer issue, reported against HBM like the other two; algorithmic bytes =
            # every preprocessed / wire / Z column of the LDE read once + the two quotient columns written (64N (P + W + 22))
            "quotient_roofline": {"kernel": "k_quotient_perm + k_quotient_limb + k_quotient_gate<kind> (vanishing-polynomial evaluation on the 8N coset)",
                                  "bound": "hbm", "achieved": q_alg / 1e9 / (mean("quotient_kernel_ms") / 1e3), "peak": hbm, "unit": "GB/s",
                                  "frac": q_alg / 1e9 / (mean("quotient_kernel_ms") / 1e3) / hbm, "algorithmic_bytes_per_step": q_alg,
                                  "traffic": NCU["quotient_dram_bytes"], "traffic_unit": "bytes per step (ncu dram__bytes_read + write, 2^20-row proof)",
                                  "avg_stage_ms": mean("quotient_kernel_ms"), "limiter": NCU["quotient_limiter"]},
            "stages_ms": dict(stages, measured_on="one proof at a time (no overlap)" if F > 1 else "the timed region"),
            "clocks": clocks,
        }
        if sharded is not None:
            line["sharded_proof"] = dict(sharded, note=f"ONE proof of the same circuit coset-sharded across the {world} GPUs "
                                                      "(trace coefficients exchanged by peer stores fused into the inverse NTT; NCCL all-gathers of subtree caps, "
                                                      "quotient values and opened rows)")
        if not args.no_cpu_baseline and world == 1:
            bits = pick_cpu_sample_bits(p2g, args, 15.0)
            dt, cores, _ = cpu_proof(p2g, args, bits, 300)
            scale = float(1 << (args.degree_bits - bits))
            line["cpu_baseline"] = {"value": 1.0 / (dt * scale), "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"one oracle (C/OpenMP port of plonky2's prover, -march=native) proof, {sample_note(args, bits)}, in "
                                              f"{dt:.2f} s, time scaled x{int(scale)} (extrapolated; `bench.py --impl reference` times a real "
                                              f"2^{args.degree_bits}-row proof)"}
        else:
            line["cpu_baseline"] = None
        emit(line)
    if sharded is None:
        for hd in handles:
            hd.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
