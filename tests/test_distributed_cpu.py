"""world_size-2 gloo test of the multi-GPU scheme on the CPU: each rank renders its spp shard with
global lane indices (here with the oracle standing in for the kernels), the film accumulators and
the gradient grids are all-reduced with the package's own helpers, and the result must equal the
single-process render up to float summation order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import media, rel_l2, small_scene


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from recursive_control_variates_b200 import parallel
    assert parallel.world() == (rank, world)
    desc = small_scene(16, 12).to_desc()
    old, new = media(8)
    mo, mn = O.OracleMedium(desc, *old), O.OracleMedium(desc, *new)
    spp = 5                                          # not divisible by the world size on purpose
    b, e = parallel.shard_spp(spp, rank, world)
    acc_o, acc_n = O.render_twostates_accum(desc, mo, mn, 11, spp, b, e, threads=1)
    gimg = np.full((12, 16, 3), 1e-3, np.float32)
    # the film weights of ALL samples: every rank contributes its own, one small all-reduce (rcv_film_weights)
    fw = torch.from_numpy(O.film_weights(desc, 99, spp, b, e))
    parallel.allreduce_sum_([fw])
    assert np.array_equal(fw.numpy(), O.film_weights(desc, 99, spp))
    gs, ga = O.vps_backward(desc, mn, 99, spp, gimg, b, e, threads=1, film_w=fw.numpy())
    # the adjoint with the DRT term shards the same way: every lane owns its secondary sampler
    ds, da = O.vps_backward(desc, mn, 99, spp, gimg, b, e, threads=1, film_w=fw.numpy(), use_drt=True)
    # the boxed scene (surfaces + area light) with the inherited prbvolpath adjoint shards the same way
    from recursive_control_variates_b200.synthetic import boxed_quads
    bdesc = small_scene(16, 12, quads=boxed_quads(), max_depth=6).to_desc()
    bm = O.OracleMedium(bdesc, *new)
    b_o, b_n = O.render_twostates_accum(bdesc, bm, bm, 11, spp, b, e, threads=1)
    bs, ba = O.prb_backward(bdesc, bm, 99, spp, gimg, b, e, threads=1, film_w=fw.numpy())
    ts = [torch.from_numpy(acc_o), torch.from_numpy(acc_n), torch.from_numpy(gs), torch.from_numpy(ga), torch.from_numpy(ds),
          torch.from_numpy(da), torch.from_numpy(b_n), torch.from_numpy(bs), torch.from_numpy(ba)]
    parallel.allreduce_sum_(ts)
    if rank == 0:
        np.savez(os.path.join(out_dir, "sharded.npz"), acc_o=ts[0].numpy(), acc_n=ts[1].numpy(), gs=ts[2].numpy(), ga=ts[3].numpy(),
                 ds=ts[4].numpy(), da=ts[5].numpy(), b_n=ts[6].numpy(), bs=ts[7].numpy(), ba=ts[8].numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_spp_sharded_two_ranks_equal_single_process(oracle, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "sharded.npz")
    desc = small_scene(16, 12).to_desc()
    old, new = media(8)
    mo, mn = oracle.OracleMedium(desc, *old), oracle.OracleMedium(desc, *new)
    acc_o, acc_n = oracle.render_twostates_accum(desc, mo, mn, 11, 5, threads=1)
    gs, ga = oracle.vps_backward(desc, mn, 99, 5, np.full((12, 16, 3), 1e-3, np.float32), threads=1)
    assert rel_l2(got["acc_o"], acc_o) < 1e-6 and rel_l2(got["acc_n"], acc_n) < 1e-6
    assert np.array_equal(got["acc_o"][..., 3], acc_o[..., 3])          # film weights: exact
    assert rel_l2(got["gs"], gs) < 1e-9 and rel_l2(got["ga"], ga) < 1e-9
    ds, da = oracle.vps_backward(desc, mn, 99, 5, np.full((12, 16, 3), 1e-3, np.float32), threads=1, use_drt=True)
    assert rel_l2(got["ds"], ds) < 1e-9 and rel_l2(got["da"], da) < 1e-9
    from recursive_control_variates_b200.synthetic import boxed_quads
    bdesc = small_scene(16, 12, quads=boxed_quads(), max_depth=6).to_desc()
    bm = oracle.OracleMedium(bdesc, *new)
    _, b_n = oracle.render_twostates_accum(bdesc, bm, bm, 11, 5, threads=1)
    bs, ba = oracle.prb_backward(bdesc, bm, 99, 5, np.full((12, 16, 3), 1e-3, np.float32), threads=1)
    assert rel_l2(got["b_n"], b_n) < 1e-6 and rel_l2(got["bs"], bs) < 1e-9 and rel_l2(got["ba"], ba) < 1e-9
