"""N-rank check of the library's exchange (run under torchrun on an N-GPU box; not collected by pytest):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_exchange_check.py
Every rank matches its round-robin share of the starting views, the ranks exchange through eg3d_points_allgather, and every rank
must hold exactly the unsharded result of the whole seed list (computed on the same GPU) — all fields, bit for bit — for the
sweep form (keyed merge) and for pipeline 3 (rank-major track blocks)."""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from edgegraph3d_b200 import lib as E, synthetic as syn, multigpu as mg
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = syn.make_scene(n_views=16, n_curves=40, seed=5, n_tracks=200)
    seeds = syn.sample_seeds(E.sample_seeds, sc, per_view=60)
    mine = np.where(seeds.view % world == rank)[0]
    fields = ("xyz", "seed", "chain_pos", "obs_off", "obs_view", "obs_poly", "obs_seg", "obs_xy")
    with E.DeviceScene(sc) as dev:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(E.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        dev.comm_create(uid.cpu().numpy().tobytes(), rank, world)
        full, _ = dev.match_seeds(seeds)
        dp, _ = dev.match_seeds(seeds.take(mine), fetch=False)
        merged, tm = dev.points_allgather(dp, mine.astype(np.int64), fetch=True)
        ok1 = all(np.array_equal(getattr(merged, f), getattr(full, f)) for f in fields) and full.n_points > 0
        tb, te = mg.track_block(sc.n_tracks, world, rank)
        full3, _ = dev.match_refpoints(0, sc.n_tracks)
        dp3, _ = dev.match_refpoints(tb, te, fetch=False)
        m3, _ = dev.points_allgather(dp3, None, fetch=True)
        ok3 = all(np.array_equal(getattr(m3, f), getattr(full3, f)) for f in fields if f != "seed") and full3.n_points > 0
        ok3 = ok3 and np.all(np.diff(m3.seed) >= 0)
        dp.free(); dp3.free()
        dev.comm_destroy()
    res = torch.tensor([int(ok1), int(ok3)], device="cuda")
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"exchange check on {world} ranks: sweep form {'IDENTICAL' if res[0] else 'DIFFERENT'} ({full.n_points} points, exchange+merge {tm['total_ms']:.3f} ms), "
              f"pipeline 3 {'IDENTICAL' if res[1] else 'DIFFERENT'} ({full3.n_points} points)")
    dist.destroy_process_group()
    sys.exit(0 if bool(res.min()) else 1)


if __name__ == "__main__":
    main()
