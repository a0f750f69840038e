"""Multi-process path on the CPU (gloo, world_size 2): the scene-level data-parallel sharding bench.py uses and the
max-over-ranks timing reduction.  Scenes never exchange data, so the only collectives are a barrier and an
all_reduce(MAX) of the elapsed time."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard_scene_seeds(rank, world, bs, step_seed0=0):
    """rank r owns scenes r*bs .. r*bs+bs-1 (DistributedSampler(shuffle=False) order) -- same rule as bench.py."""
    return [step_seed0 + rank * bs + s for s in range(bs)]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from focalformer3d_b200.config import load_config, default_config_path, scaled_model_cfg
    from focalformer3d_b200.synth import make_state_dict, synth_points
    from oracle.detector import build_oracle
    cfg = scaled_model_cfg(load_config(default_config_path())["model"], bev=16, num_proposals=8)
    oracle = build_oracle(cfg)
    oracle.load_state_dict(make_state_dict(cfg, 0), strict=True)
    seeds = shard_scene_seeds(rank, world, bs=1)
    pts = [torch.from_numpy(synth_points(3000, cfg["pts_voxel_layer"]["point_cloud_range"], seed=s)) for s in seeds]
    dist.barrier()
    out, _ = oracle.forward_raw(pts)
    elapsed = torch.tensor([float(rank + 1)])                       # stand-in for the per-rank CUDA-event time
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, (seeds, out["center"].sum().item()))
    dist.barrier()
    if rank == 0:
        q.put((elapsed.item(), gathered))
    dist.destroy_process_group()


def test_two_rank_scene_sharding_matches_single_process():
    sys.path.insert(0, ROOT)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    elapsed, gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert elapsed == 2.0                                           # MAX over ranks
    seeds = [g[0] for g in gathered]
    assert seeds == [[0], [1]] and sorted(sum(seeds, [])) == [0, 1]  # disjoint, complete cover of the global batch
    # a rank's result equals the single-process result for the same scene (no cross-scene coupling)
    from focalformer3d_b200.config import load_config, default_config_path, scaled_model_cfg
    from focalformer3d_b200.synth import make_state_dict, synth_points
    from oracle.detector import build_oracle
    cfg = scaled_model_cfg(load_config(default_config_path())["model"], bev=16, num_proposals=8)
    oracle = build_oracle(cfg)
    oracle.load_state_dict(make_state_dict(cfg, 0), strict=True)
    for (sd, val) in gathered:
        pts = [torch.from_numpy(synth_points(3000, cfg["pts_voxel_layer"]["point_cloud_range"], seed=sd[0]))]
        out, _ = oracle.forward_raw(pts)
        assert abs(out["center"].sum().item() - val) < 1e-3
