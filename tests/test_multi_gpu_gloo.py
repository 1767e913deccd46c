"""`not gpu`: the N>1 path (image sharding + the single all-gather of person records) with
world_size 2 over gloo on CPU.  The records are synthetic; the kernels are not involved."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import pkg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_records(n_images, max_persons, seed):
    native = pkg("_native")
    rs = np.random.RandomState(seed)
    headers = np.zeros(n_images, native.HEADER_DTYPE)
    persons = np.zeros((n_images, max_persons), native.PERSON_DTYPE)
    for i in range(n_images):
        k = rs.randint(0, max_persons + 1)
        headers[i] = (rs.randint(0, 200), k, 0, rs.randint(0, 300))
        persons["score"][i, :k] = rs.uniform(1, 30, k)
        persons["count"][i, :k] = rs.randint(3, 19, k)
        persons["peak_id"][i, :k] = rs.randint(-1, 150, (k, 18))
        persons["x"][i, :k] = rs.randint(0, 576, (k, 18))
        persons["y"][i, :k] = rs.randint(0, 320, (k, 18))
    return headers, persons


def _worker(rank, world, port, n_images, max_persons, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mg = pkg("multi_gpu")
    headers, persons = _fake_records(n_images, max_persons, seed=123)      # the global truth
    a, b = mg.shard_range(n_images, world, rank)
    per = -(-n_images // world)
    local = np.zeros((per, mg.record_bytes(max_persons)), np.uint8)
    local[:b - a] = mg.pack_records(headers[a:b], persons[a:b])
    gathered = mg.all_gather_records(torch.from_numpy(local.reshape(-1)), per)
    gh, gp = mg.gathered_to_global(gathered.numpy(), n_images, world, max_persons)
    ok = np.array_equal(gh, headers) and np.array_equal(gp, persons)
    with open(os.path.join(result_dir, "rank%d" % rank), "w") as f:
        f.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [8, 7, 1])
def test_shard_and_all_gather_world2(tmp_path, n_images):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_images, 6, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "rank%d" % r)).read() == "ok"


def test_shard_range_partitions_every_image_once():
    mg = pkg("multi_gpu")
    for n in (1, 2, 7, 32, 255, 256):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                a, b = mg.shard_range(n, world, r)
                assert 0 <= a <= b <= n
                seen += list(range(a, b))
            assert seen == list(range(n))


def test_pack_unpack_roundtrip():
    mg = pkg("multi_gpu")
    h, p = _fake_records(5, 4, seed=1)
    h2, p2 = mg.unpack_records(mg.pack_records(h, p), 4)
    assert np.array_equal(h, h2) and np.array_equal(p, p2)
