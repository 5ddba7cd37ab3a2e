"""In-library NCCL (p2g_circuit_create_sharded_nccl): one rank per GPU of this box, the ranks driven by threads of this process.
Needs at least two GPUs (NCCL refuses two ranks on one device), so the single-GPU test run skips it; `gpurun --gpus 2` runs it.
The proof must equal the single-GPU proof and the oracle's, and every exchange must have gone through the library's communicator."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpus(p2g):
    return p2g.lib.lib().p2g_device_count()


@pytest.mark.parametrize("degree_bits,workload,hasher", [(12, "ecdsa", "keccak25"), (13, "range", "poseidon"), (16, "assert_zero", "keccak25")])
def test_nccl_sharded_proof_matches_single_gpu_and_oracle(p2g, corc, degree_bits, workload, hasher):
    n = _ngpus(p2g)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    from helpers import oracle_prove_and_verify
    world = 1 << (min(n, 8).bit_length() - 1)
    cfg = p2g.CircuitConfig.wide_ecc_config(hasher=hasher)
    sc = p2g.synth.SyntheticCircuit(degree_bits, workload, config=cfg, num_public_inputs=2, seed=5000 + degree_bits)
    ref_bytes, _ = oracle_prove_and_verify(corc, sc)
    with p2g.CircuitData(sc.common, sc.constants_sigmas, device=0) as single:
        assert single.prove(sc.wires, sc.public_inputs).to_bytes() == ref_bytes
    uid = p2g.sharding.nccl_unique_id()
    outs, errs = [None] * world, []

    def rank_main(r):
        try:
            grp = p2g.sharding.NcclGroup(r, world, r, uid)
            with p2g.CircuitData(sc.common, sc.constants_sigmas, device=r, shard=grp) as d:
                info = d.read(p2g.lib.BUF_SHARD_INFO)
                assert (int(info[0]), int(info[1]), int(info[4])) == (r, world, 1)
                a = d.prove(sc.wires, sc.public_inputs).to_bytes()
                b = d.prove(sc.wires, sc.public_inputs, timings=False).to_bytes()
                assert int(d.read(p2g.lib.BUF_SHARD_INFO)[5]) > int(info[5])
                outs[r] = (a, b)
        except BaseException as e:  # noqa: BLE001
            errs.append(e)
    ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        raise errs[0]
    for a, b in outs:
        assert a == ref_bytes and b == ref_bytes


def test_nccl_unique_id_and_bad_arguments(p2g):
    import ctypes as C
    uid = p2g.sharding.nccl_unique_id()
    assert len(uid) == 128 and uid != p2g.sharding.nccl_unique_id()
    sc = p2g.synth.SyntheticCircuit(5, "assert_zero", seed=1)
    desc, keep = sc.common.fill_desc(sc.constants_sigmas)
    h = C.c_void_p()
    L = p2g.lib.lib()
    assert L.p2g_circuit_create_sharded_nccl(C.byref(desc), 0, 0, 2, None, C.byref(h)) == p2g.lib.P2G_EBADARG
    assert L.p2g_circuit_create_sharded_nccl(C.byref(desc), 0, 0, 3, uid, C.byref(h)) == p2g.lib.P2G_EBADARG      # not a power of two
    assert L.p2g_circuit_create_sharded_nccl(C.byref(desc), 0, 2, 2, uid, C.byref(h)) == p2g.lib.P2G_EBADARG      # rank out of range
    # world == 1 needs no communicator: it is the plain single-GPU handle
    assert L.p2g_circuit_create_sharded_nccl(C.byref(desc), 0, 0, 1, None, C.byref(h)) == 0
    L.p2g_circuit_destroy(h)
    del keep
