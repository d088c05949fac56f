"""world_size-2 gloo check of the scene-sharding host logic: each rank super-resolves its band of
patch-grid rows, the stitched stripes are all-gathered, and every rank ends with the single-process mosaic.
The kernels are replaced by their torch statements (tests/opref.RefOps) so this runs without a GPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import lfsr_b200
    from opref import RefOps
    from oracle import weights
    name, scale, A = "LF_InterNet", 4, 5
    net = lfsr_b200.load_net(name, A, scale).eval()
    net.load_state_dict(weights.make_state_dict(name, scale, 1234))
    net.set_backend(RefOps())
    lr = torch.from_numpy(np.random.RandomState(5).random_sample((A * 40, A * 24)).astype(np.float32))
    sr = lfsr_b200.scene.super_resolve_scene(net, lr, A, scale, 32, 16, minibatch=2, ops=RefOps())
    np.save(os.path.join(out_dir, f"sr_rank{rank}.npy"), sr.numpy())
    lo, hi = lfsr_b200.scene.shard_rows(3, world, rank)
    np.save(os.path.join(out_dir, f"rows_rank{rank}.npy"), np.array([lo, hi]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_rows_partition():
    import lfsr_b200
    for num_u in (1, 3, 7, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [lfsr_b200.scene.shard_rows(num_u, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == num_u
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(900)
def test_two_rank_scene_equals_single_process(tmp_path):
    import lfsr_b200
    from opref import RefOps
    from oracle import lf_oracle, nets as onets, weights
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "sr_rank0.npy")
    b = np.load(tmp_path / "sr_rank1.npy")
    assert np.array_equal(a, b)
    r0, r1 = np.load(tmp_path / "rows_rank0.npy"), np.load(tmp_path / "rows_rank1.npy")
    assert r0[0] == 0 and r0[1] == r1[0] and r1[1] == 3
    # single-process oracle of the same scene
    name, scale, A = "LF_InterNet", 4, 5
    sd = weights.make_state_dict(name, scale, 1234)
    lr = np.random.RandomState(5).random_sample((A * 40, A * 24)).astype(np.float32)
    sub = lf_oracle.lfdivide(lr, A, 32, 16)
    nu, nv = sub.shape[:2]
    y = onets.forward(name, torch.from_numpy(sub.reshape(nu * nv, 1, 160, 160)), sd, A, scale).numpy()
    want = lf_oracle.to_sai(lf_oracle.lfintegrate(y.reshape(nu, nv, 640, 640), A, 128, 64, 160, 96))
    assert np.abs(a - want).max() <= 2e-5
