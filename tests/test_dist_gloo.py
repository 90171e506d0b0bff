"""world_size-2 gloo test (CPU) of the multi-rank protocol: chains sharded by contiguous blocks,
keys from the global chain index, diagnostics as ONE all-reduce(sum) of sufficient statistics
followed by the native host finalisation."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import diagnostics as OD
from oracle import prng as P
from oracle import samplers as S
from oracle import targets as T
from tests.test_cabi_and_host import _ar1, partial_stats


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from geomjax_b200 import _native as N
        lib = N.lib()
        Tn, Cn, D = 120, 6, 3
        x = _ar1(Tn, Cn, D, 0.6, seed=5)                      # the GLOBAL sample array
        lo, hi = rank * Cn // world, (rank + 1) * Cn // world   # this rank's contiguous chain block
        stats, acov = partial_stats(x[:, lo:hi])
        stats_t, acov_t = torch.from_numpy(stats), torch.from_numpy(np.ascontiguousarray(acov[:32]))
        dist.all_reduce(stats_t)
        dist.all_reduce(acov_t)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        rhat, ess, tr = np.empty(D), np.empty(D), np.zeros(D, np.uint8)
        s_np, a_np = stats_t.numpy(), acov_t.numpy()
        assert lib.gb200_rhat_finalize(dp(s_np), Tn, D, dp(rhat)) == 0
        assert lib.gb200_ess_finalize(dp(a_np), dp(s_np), Tn, Cn, D, 32, dp(ess), tr.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
        np.testing.assert_allclose(rhat, OD.potential_scale_reduction(x, 1, 0), rtol=1e-10)
        np.testing.assert_allclose(ess[tr == 0], OD.effective_sample_size(x, 1, 0)[tr == 0], rtol=1e-8)
        assert (tr == 0).all()

        # sharded transitions: the rank's shard of the global key tree / chain set gives the same
        # chains as the single-process run (oracle stands in for the kernel on CPU)
        tgt = T.NealFunnel(4)
        Ctot = 8
        root = P.key(3)
        q0 = np.ones((Ctot, 4), np.float32)
        full, _ = S.lmcmonge_step(S.chain_keys(root, 10, 2, Ctot), S.lmcmonge_init(q0, tgt), tgt, 0.01,
                                  np.ones(4, np.float32), 3)
        lo, hi = rank * Ctot // world, (rank + 1) * Ctot // world
        mine, _ = S.lmcmonge_step(S.chain_keys(root, 10, 2, Ctot, np.arange(lo, hi)),
                                  S.lmcmonge_init(q0[lo:hi], tgt), tgt, 0.01, np.ones(4, np.float32), 3)
        np.testing.assert_array_equal(mine.position, full.position[lo:hi])
        gathered = [torch.zeros((hi - lo, 4)) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(mine.position))
        np.testing.assert_array_equal(torch.cat(gathered).numpy(), full.position)
        open(os.path.join(out, f"rank{rank}.ok"), "w").write("1")
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_diagnostics_and_keys(tmp_path):
    # results come back through files: a forked multiprocessing.Manager left the parent's BLAS thread
    # pool in a state where later (larger) NumPy products in the same pytest process dead-locked
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["rank0.ok", "rank1.ok"]
