"""CPU tests of the host-side logic around the C ABI (no compute call needs a GPU):

* the shared library loads and exports every symbol include/dqmc_gpu.h declares;
* the host-only entry points (exchange probability, serial ladder walk, RNG stream) agree with the
  oracle;
* the multi-rank replica-exchange path (ReplicaExchangeLadder + one all-gather per exchange step)
  on 2 ranks over gloo reproduces the single-process serial walk of DetQMCPT::replicaExchangeStep
  (detqmcpt.h:1031-1079) as restated by the oracle.
"""
import os
import re
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "dqmc_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dqmc_[a-z0-9_]+)\s*\(", text)) - {"dqmc_rng_fill_fn"})


def test_library_exports_every_declared_symbol():
    import ctypes
    from detqmc_b200.lib import LIB_PATH, SYMBOLS, load_library
    lib = load_library()
    declared = _header_symbols()
    assert len(declared) >= 40
    raw = ctypes.CDLL(LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), "libdqmc_b200.so does not export " + name
    bound = {s[0] for s in SYMBOLS}
    assert set(declared) <= bound, "ctypes bindings miss: %s" % sorted(set(declared) - bound)
    assert lib.dqmc_profile_name(0).decode() == "cb_mult"


def test_library_is_built_for_sm_100a_with_the_hardware_paths_the_design_claims():
    """The shipped library is sm_100a machine code and contains what DESIGN.md says the hot kernels use: FP64
    tensor-core MMAs (DMMA), cp.async (LDGSTS), bulk copies through the copy engine with mbarrier completion (UBLKCP,
    SYNCS) and thread-block-cluster barriers (UCGABAR).  Skipped where the CUDA binary tools are not installed."""
    import shutil
    import subprocess
    from detqmc_b200.lib import LIB_PATH
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    elf = subprocess.run(["cuobjdump", "-lelf", LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"\.(sm_[0-9a-z]+)\.cubin", elf))
    assert archs == {"sm_100a"}, archs
    sass = subprocess.run(["cuobjdump", "-sass", LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("DMMA.8x8x4", "LDGSTS", "UBLKCP", "SYNCS", "UCGABAR"):
        assert mnemonic in sass, mnemonic


def test_create_without_gpu_fails_loudly():
    """No CPU fallback: without a CUDA device dqmc_create must return an error, not a context that works."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from detqmc_b200 import DetSDWBatch, DqmcError
    from dqmc_oracle import SdwParams
    with pytest.raises(DqmcError):
        DetSDWBatch(SdwParams(), n_replicas=1, device=0)


def test_exchange_probability_and_walk_match_oracle():
    from detqmc_b200 import load_library
    from detqmc_b200.sdw import exchange_walk
    from dqmc_oracle import exchange_probability, replica_exchange_walk
    lib = load_library()
    gen = np.random.default_rng(11)
    for _ in range(200):
        p1, p2 = gen.uniform(-2, 1, 2)
        a1, a2 = gen.uniform(0, 50, 2)
        ref = exchange_probability(p1, a1, p2, a2)                   # libm exp vs numpy exp: last-bit freedom
        assert abs(lib.dqmc_exchange_probability(p1, a1, p2, a2) - ref) <= 4e-16 * ref
    P = 16
    ladder = np.linspace(-1.9, 0.4, P)
    par_proc = np.arange(P, dtype=np.int32)
    proc_par = np.arange(P, dtype=np.int32)
    o_pp, o_pr = list(range(P)), list(range(P))
    for it in range(40):
        actions = gen.uniform(10, 12, P)
        uniforms = gen.uniform(0, 1, P - 1)
        pos = [0]

        def rand01():
            pos[0] += 1
            return uniforms[pos[0] - 1]
        o_pp, o_pr, log = replica_exchange_walk(ladder, o_pp, o_pr, actions, rand01)
        used, swapped = exchange_walk(ladder, par_proc, proc_par, actions, uniforms)
        assert used == pos[0]
        assert list(par_proc) == o_pp and list(proc_par) == o_pr
        assert [bool(s) for s in swapped] == [a for _, _, a in log]


def test_rng_stream_matches_oracle():
    """rng_stream.cpp (dSFMT-19937 restated for the product) against the oracle's generator and the
    reference's seed scramble KAT (rngwrapper.cpp:43: seed 1020304050, index 1 -> 37767)."""
    import ctypes
    from dsfmt_oracle import RngOracle
    from detqmc_b200 import load_library
    lib = load_library()
    for seed, idx in [(1020304050, 1), (1020304050, 7), (5, 3)]:
        n = 5000
        out = np.zeros(n)
        assert lib.dqmc_rng_stream_sample(seed, idx, n, out.ctypes.data_as(ctypes.c_void_p)) == 0
        o = RngOracle(seed, idx)
        ref = np.array([o.rand01() for _ in range(n)])
        assert np.array_equal(out, ref)


# ---------------------------------------------------------------------------------------------
# 2 ranks over gloo
# ---------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ladder_rank(rank, world, port, n_steps, P, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from detqmc_b200.sdw import ReplicaExchangeLadder, CTRL_WORDS
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_local = P // world
    ladder = np.linspace(-1.9, 0.4, P)
    lad = ReplicaExchangeLadder(ladder, n_local, rank, world)
    gen = np.random.default_rng(100 + rank)          # each rank only knows ITS replicas' data
    ugen = np.random.default_rng(7)                  # replica 0's stream lives on rank 0
    blobs = np.zeros((n_local, CTRL_WORDS))
    blobs[:, 0] = 1000.0 * rank + np.arange(n_local)  # tag: "phiDelta" identifies the owner
    r_hist, tag_hist = [], []
    for step in range(n_steps):
        actions = gen.uniform(10, 12, n_local)
        uniforms = ugen.uniform(0, 1, P - 1) if rank == 0 else np.zeros(P - 1)
        payload = torch.from_numpy(np.concatenate([actions, uniforms, blobs.reshape(-1)]))
        assert payload.numel() == lad.payload_len
        gathered = torch.zeros(world * lad.payload_len, dtype=torch.float64)
        dist.all_gather_into_tensor(gathered, payload)
        r_new, ctrl_new, used = lad.walk(gathered.numpy())
        blobs = ctrl_new.reshape(n_local, CTRL_WORDS).copy()
        assert (used > 0) <= (rank == 0)
        r_hist.append(r_new.copy())
        tag_hist.append(blobs[:, 0].copy())
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), r=np.array(r_hist), tag=np.array(tag_hist),
             accepted=lad.accepted)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_replica_exchange_two_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp
    from dqmc_oracle import replica_exchange_walk
    from detqmc_b200.sdw import CTRL_WORDS     # noqa: F401
    world, P, n_steps = 2, 8, 25
    port = _free_port()
    mp.spawn(_ladder_rank, args=(world, port, n_steps, P, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    # serial restatement with the same inputs
    n_local = P // world
    gens = [np.random.default_rng(100 + r) for r in range(world)]
    ugen = np.random.default_rng(7)
    ladder = np.linspace(-1.9, 0.4, P)
    pp, pr = list(range(P)), list(range(P))
    tags = np.concatenate([1000.0 * r + np.arange(n_local) for r in range(world)])   # tag held by replica pi
    n_acc = np.zeros(P - 1, dtype=np.int64)
    for step in range(n_steps):
        actions = np.concatenate([g.uniform(10, 12, n_local) for g in gens])
        uniforms = ugen.uniform(0, 1, P - 1)
        pos = [0]

        def rand01():
            pos[0] += 1
            return uniforms[pos[0] - 1]
        old_pp = list(pp)
        pp, pr, log = replica_exchange_walk(ladder, pp, pr, actions, rand01)
        n_acc += np.array([a for _, _, a in log], dtype=np.int64)
        # control data follow the parameter: replica pi takes the blob of the previous holder of its new parameter
        tags = np.array([tags[old_pp[pr[pi]]] for pi in range(P)])
        for r in range(world):
            lo = r * n_local
            assert np.array_equal(res[r]["r"][step], ladder[np.array(pr[lo:lo + n_local])])
            assert np.array_equal(res[r]["tag"][step], tags[lo:lo + n_local])
    for r in range(world):
        assert np.array_equal(res[r]["accepted"][:P - 1], n_acc)
    assert n_acc.sum() > 0
