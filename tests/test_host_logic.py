"""CPU-only: the host mirror of plonky2's CommonCircuitData (gate ordering, selector groups, derived counts, FRI schedule,
gate constructors) against the oracle's independent derivation and against what the golden proofs imply."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def test_common_data_matches_oracle_for_every_workload(p2g):
    from helpers import oracle_cd
    for hasher in ("keccak25", "poseidon"):
        for wl in ("assert_zero", "sha256", "ecdsa", "range", "all_gates"):
            for bits in (3, 8, 16, 20, 22):
                cfg = p2g.CircuitConfig.wide_ecc_config(hasher=hasher)
                gates = [g for g, _ in p2g.synth.gate_mix(wl, cfg)] + [p2g.Gate.noop(), p2g.Gate.public_input(), p2g.Gate.poseidon()]
                com = p2g.CommonCircuitData(cfg, bits, gates, num_public_inputs=4)
                oracle_cd(com)      # asserts gate order, selector indices/groups, constants, constraints, partial products, k_is


def test_fri_schedule_is_the_surveyed_one(p2g):
    cfg = p2g.CircuitConfig.wide_ecc_config()
    sched = {n: p2g.CommonCircuitData(cfg, n, [p2g.Gate.noop()]).reduction_arity_bits for n in (3, 16, 18, 20, 22)}
    assert sched == {3: [], 16: [4, 4, 4], 18: [4, 4, 4, 4], 20: [4, 4, 4, 4], 22: [4, 4, 4, 4, 4]}   # SURVEY 8a, FRI schedule


def test_gate_constructors_follow_the_reference_formulas(p2g):
    cfg = p2g.CircuitConfig.wide_ecc_config()
    G = p2g.Gate
    assert G.u32_arithmetic(cfg).params[0] == 6          # arithmetic_u32.rs:40-43: min(234 / 38, 80 / 6)
    assert G.u32_subtraction(cfg).params[0] == 11        # subtraction_u32.rs:38-42: min(234 / 21, 80 / 5)
    assert G.u32_add_many(cfg, 2).params[:2] == (2, 10)  # add_many_u32.rs:43-48
    assert G.u32_add_many(cfg, 16).params[:2] == (16, 4)
    assert G.arithmetic(cfg).params[0] == 20
    assert G.random_access(cfg, 4).params[:3] == (4, 4, 2)
    assert G.u32_arithmetic(cfg).num_constraints == 36 * 6 and G.comparison(32, 16).num_constraints == 6 + 5 * 16 + 2
    assert G.poseidon().num_constraints == 123 and G.poseidon().degree == 7


def test_golden_circuit_shape_recovered_by_the_mirror(p2g):
    """The two committed proofs were made with 135 wires, degree 2^3; the mirror reproduces the oracle's CommonData for them."""
    from oracle.pyref import golden
    for name in ("basic_if", "basic_div"):
        cd = golden.recover(name)["cd"]
        C = p2g.circuit
        cfg = C.CircuitConfig(num_wires=cd.num_wires, num_routed_wires=cd.num_routed, hasher=cd.hasher)
        com = C.CommonCircuitData(cfg, cd.degree_bits, [C.Gate(g.kind, tuple(g.params)) for g in cd.gates], cd.num_public_inputs)
        assert com.selector_indices == cd.selector_indices and com.groups == cd.groups
        assert com.num_constants == cd.num_constants and com.num_gate_constraints == cd.num_gate_constraints
        assert com.reduction_arity_bits == cd.arity_bits == []


def test_shapes_are_checked_before_cuda(p2g):
    sc = p2g.synth.SyntheticCircuit(4, "assert_zero", seed=4)
    with pytest.raises(ValueError):
        sc.common.fill_desc(sc.constants_sigmas[:-1])
