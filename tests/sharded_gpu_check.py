"""torchrun --nproc-per-node N tests/sharded_gpu_check.py [workload]
Every rank runs its band of the sharded exchange (NCCL halo) AND the whole unsharded exchange on
its own GPU, then compares its band of every output bit for bit."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    dccm = importlib.import_module("dennou-ccm_b200")
    syn = importlib.import_module("dennou-ccm_b200.synthetic")
    X = importlib.import_module("dennou-ccm_b200.exchange")
    sh = importlib.import_module("dennou-ccm_b200.sharding")
    sys.argv += ["T106_1deg"]
    import bench
    A, O, S, K, nc, M = bench.make_grids(dccm, sys.argv[1])
    full = X.SurfaceExchange(A, O, S, K, nc, 1, device=dev)
    full.set_inputs(syn.column_inputs(torch, A, K, nc, dev=dev),
                    {k: v[None] for k, v in syn.atm_surface_fields(torch, A, dev=dev).items()},
                    {k: v[None] for k, v in syn.ocn_surface_fields(torch, O, dev=dev).items()})
    full.step()
    halo = os.environ.get("DCCM_HALO", "peer")           # peer | nccl | allgather
    cls = sh.PeerShardedExchange if halo == "peer" else sh.ShardedExchange
    kw = {"halo": "allgather"} if halo == "allgather" else {}
    if halo == "peer" and os.environ.get("DCCM_SYNC"):
        kw["sync"] = os.environ["DCCM_SYNC"]
    ex = cls(A, O, S, K, nc, 1, rank=rank, world=world, dist=dist, device=dev, **kw)
    (a0, a1), (o0, o1) = ex.plan.bands["A"][rank], ex.plan.bands["O"][rank]
    ex.set_inputs(syn.column_inputs(torch, A, K, nc, a0, a1, dev=dev),
                  {k: v[None] for k, v in syn.atm_surface_fields(torch, A, a0, a1, dev=dev).items()},
                  {k: v[None] for k, v in syn.ocn_surface_fields(torch, O, o0, o1, dev=dev).items()})
    for _ in range(6):   # repeatedly: buffer reuse across exchanges is ordered by the two handshakes
        ex.step()
    torch.cuda.synchronize()
    bad = []
    ca, co = slice(a0 * A.im, a1 * A.im), slice(o0 * O.im, o1 * O.im)
    if not torch.equal(ex.a_recv, full.a_recv[:, ca]): bad.append("a_recv")
    if not torch.equal(ex.o_recv, full.o_recv[:, co]): bad.append("o_recv")
    for k in ("DUDt", "DVDt", "DTempDt"):
        if not torch.equal(ex.tend[k], full.tend[k][:, ca]): bad.append(k)
    if not torch.equal(ex.tend["DQMixDt"], full.tend["DQMixDt"][:, :, ca]): bad.append("DQMixDt")
    t = torch.tensor([len(bad)], device=dev)
    dist.all_reduce(t)
    form = {1: "staged", 0: "direct"}.get(ex.sfc_last_form(), "?")
    print(f"rank {rank}/{world} [{cls.__name__} halo={halo}, fused surface kernel: {form}]: bands A{(a0, a1)} O{(o0, o1)} mismatches {bad}", flush=True)
    dist.destroy_process_group()
    sys.exit(1 if int(t.item()) else 0)


if __name__ == "__main__":
    main()
