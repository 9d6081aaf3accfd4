"""World-size-2 gloo run of the multi-GPU split: two CPU processes each accumulate their own subframe range
with the oracle, merge with one sum-reduce of the moment buffers, and rank 0 compares against the sequential
single-process Welford accumulation."""
import lmdb_compat
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent

from deepestscatter_b200.multigpu import subframe_range  # noqa: E402


def test_subframe_range_partitions_exactly():
    for total in (0, 1, 7, 64, 1024, 8192):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                off, cnt = subframe_range(r, world, total)
                seen += list(range(off + 1, off + cnt + 1))
            assert seen == list(range(1, total + 1))
    with pytest.raises(ValueError):
        subframe_range(2, 2, 10)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out_path):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as ol
    from deepestscatter_b200 import multigpu as mg

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h = 24, 14
    o = ol.Oracle()
    o.volume_synth(32, 0, 1234)
    o.scene_set(7000.0, (-0.586, -0.766, -0.271))
    o.bake()
    cam = ol.camera_look_at(aspect=w / h)
    off, cnt = mg.subframe_range(rank, world, total)
    prog = np.zeros((h, w, 4), dtype=np.float32)
    var = np.zeros((h, w, 4), dtype=np.float32)
    for k in range(cnt):  # global stream id off+k+1, local Welford weight 1/(k+1)
        fr = o.render_frame(cam, w, h, ol.MODE_ALL, off + k + 1)
        ol.lib().orc_update_frame_result(fr.reshape(-1), prog.reshape(-1), var.reshape(-1), fr.size, k + 1)
    m = mg.export_moments(torch.from_numpy(prog), torch.from_numpy(var), cnt)
    mg.reduce_moments(m, dst=0)
    if rank == 0:
        mean, m2 = mg.import_moments(m, total)
        np.savez(out_path, mean=mean.numpy(), m2=m2.numpy())
    dist.destroy_process_group()


def test_two_rank_gloo_merge_matches_sequential_accumulation(tmp_path):
    import torch.multiprocessing as mp

    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as ol

    total, world = 9, 2
    out = tmp_path / "merged.npz"
    mp.spawn(_worker, args=(world, _free_port(), total, str(out)), nprocs=world, join=True)
    got = np.load(out)
    w, h = 24, 14
    o = ol.Oracle()
    o.volume_synth(32, 0, 1234)
    o.scene_set(7000.0, (-0.586, -0.766, -0.271))
    o.bake()
    cam = ol.camera_look_at(aspect=w / h)
    prog, var = o.render_accumulate(cam, w, h, ol.MODE_ALL, 1, total)
    assert np.allclose(got["mean"], prog, rtol=3e-6, atol=1e-7)
    scale = max(1.0, float(var.max()))
    assert np.allclose(got["m2"], var, rtol=2e-4, atol=1e-6 * scale)
    assert prog[..., 0].max() > 0


# ---------------------------------------------------------------- dataset generation: scenes sharded over ranks

def _scene_worker(rank, world, port, scenes, out_dir):
    """One rank of a sharded dataset run on CPU: draws scenes from the shared counter, writes its records (values derived
    from the scene id instead of a GPU) into its own LMDB shard."""
    import time

    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    import deepestscatter_b200 as ds
    from deepestscatter_b200 import multigpu as mg

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    queue = mg.SceneQueue(dist.distributed_c10d._get_default_store(), scenes)
    batch = 8
    mine = []
    with ds.Dataset(Path(out_dir) / f"Train.rank{rank}.lmdb") as store:
        for sid in queue:
            mine.append(sid)
            time.sleep(0.01 * (1 + sid % 3))  # uneven scene costs
            store.append_scene_setup(sid, f"cloud{sid}.npy", 1000.0 + sid, (0.0, -1.0, 0.0))
            store.append_results(sid * batch, np.full(batch, 0.5 + sid, np.float32), np.ones(batch, np.uint8))
    (Path(out_dir) / f"scenes.rank{rank}.txt").write_text(" ".join(map(str, mine)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_scene_queue_and_shard_merge(tmp_path):
    import torch.multiprocessing as mp

    sys.path.insert(0, str(ROOT))
    import deepestscatter_b200 as ds
    from deepestscatter_b200.multigpu import static_scenes

    scenes = [0, 1, 2, 3, 4, 7, 9]  # a resumed run: scenes 5, 6, 8 are already done
    world = 2
    mp.spawn(_scene_worker, args=(world, _free_port(), scenes, str(tmp_path)), nprocs=world, join=True)
    taken = [list(map(int, (tmp_path / f"scenes.rank{r}.txt").read_text().split())) for r in range(world)]
    assert sorted(taken[0] + taken[1]) == scenes and not set(taken[0]) & set(taken[1])
    assert taken[0] and taken[1]  # both ranks worked
    merged = tmp_path / "Train.lmdb"
    with ds.Dataset(merged) as out:
        for r in range(world):
            out.merge(str(tmp_path / f"Train.rank{r}.lmdb"))
        assert out.count("SceneSetup") == len(scenes) and out.count("Result") == 8 * len(scenes)
        for sid in scenes:
            assert out.get("Result", sid * 8 + 3) == ds.record_result(0.5 + sid, True)
            assert out.get("SceneSetup", sid) == ds.record_scene_setup(f"cloud{sid}.npy", 1000.0 + sid, (0.0, -1.0, 0.0))
    assert lmdb_compat.check(str(merged))["pages_leaked"] == 0
    assert static_scenes(0, 2, scenes) == [0, 2, 4, 9] and static_scenes(1, 2, scenes) == [1, 3, 7]
