"""N > 1 host-side logic on the CPU (gloo, world_size 2): the image sharding of BASELINE config 5 (image i -> rank
i mod N, every prompt answered by the rank that holds its image) and the one exchange of the job, the gather of
per-prompt IoU scores.  The data path itself has no collective."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dlimgedit_b200 import sharding  # noqa: E402


def fake_score(image: int, prompt: int) -> float:
    """Stand-in for "encode the image and answer the prompt": a deterministic score per (image, prompt)."""
    return float(((image * 131 + prompt * 17) % 1000) / 1000.0)


def _worker(rank, world, port, n_images, prompts_per_image, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_indices(n_images, rank, world)
    # stand-in for "encode my images and answer their prompts": a deterministic score per (image, prompt)
    local = np.array([[fake_score(i, p) for p in range(prompts_per_image)] for i in mine], np.float32)
    full = sharding.gather_scores(local, n_images, rank, world)
    if rank == 0:
        np.save(os.path.join(out_dir, "scores.npy"), full)
    dist.destroy_process_group()


def test_shard_indices_partition():
    for n in (0, 1, 7, 4096):
        for world in (1, 2, 4, 8):
            parts = [sharding.shard_indices(n, r, world) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert all(all(i % world == r for i in p) for r, p in enumerate(parts))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


@pytest.mark.parametrize("n_images", [10, 7])
def test_two_rank_gather(tmp_path, n_images):
    world, prompts = 2, 16
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_images, prompts, str(tmp_path)), nprocs=world, join=True)
    full = np.load(tmp_path / "scores.npy")
    ref = np.array([[fake_score(i, p) for p in range(prompts)] for i in range(n_images)], np.float32)
    assert full.shape == (n_images, prompts) and np.array_equal(full, ref)
