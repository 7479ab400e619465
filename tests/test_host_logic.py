"""CPU: host-side mirrors of the reference interface (registry, scheduler, partition, sharding)."""

import os

import numpy as np
import pytest
import torch

import nova_pointcloud_b200 as nb
from oracle import head as OH
from oracle import partition as OP
from oracle import scheduler as OS


def test_registry_semantics_and_names():
    r = nb.Registry("things")
    r.register("a", lambda x, k=1: (x, k), k=5)

    @r.register(["b", "c"], k=7)
    def f(x, k=0):
        return x + k

    assert r.get("a")(2) == (2, 5) and r.get("b")(1) == 8 and r.get("c")(1) == 8
    assert r.has("a") and r.try_get("zz") is None and r.get(None) is None
    with pytest.raises(KeyError):
        r.get("zz")
    assert r.get("zz", default=3) == 3
    for w in (768, 1024, 1536):
        assert nb.POINT_CLOUD_DECODERS.has(f"mlp_d6w{w}") and nb.IMAGE_DECODERS.has(f"mlp_d6w{w}")
    assert nb.IMAGE_DECODERS.has("mlp_d3w1280")
    h = nb.IMAGE_DECODERS.get("mlp_d3w1280")(patch_size=2, image_dim=4, cond_dim=1024)
    assert (h.depth, h.embed_dim, h.cond_dim, h.token_dim) == (3, 1280, 1024, 16)
    h = nb.POINT_CLOUD_DECODERS.get("mlp_d6w768")(1, 768)  # reference factories ignore patch_size
    assert (h.depth, h.embed_dim, h.token_dim) == (6, 768, 16)


def test_module_state_dict_contract_and_init_matches_reference_order(golden_dir):
    torch.manual_seed(101)
    head = nb.DiffusionMLP(2, 128, 96, patch_size=1, image_dim=3)
    d = np.load(os.path.join(golden_dir, "head_p1.npz"))
    sd = head.state_dict()
    ref_keys = [k[3:] for k in d.files if k.startswith("w::")]
    assert list(sd.keys()) == ref_keys and len(sd) == 14 + 8 * 2
    for k in ref_keys:  # same construction order => the reference's own random init
        assert np.array_equal(sd[k].numpy(), d["w::" + k]), k
    assert head.blocks[0].mlp_checkpointing is False


def test_module_loads_reference_state_dict(golden_dir):
    d = np.load(os.path.join(golden_dir, "head_p2.npz"))
    head = nb.DiffusionMLP(1, 64, 64, patch_size=2, image_dim=4)
    missing = head.load_state_dict({k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("w::")})
    assert not missing.missing_keys and not missing.unexpected_keys


def test_patchify_matches_oracle():
    pe = nb.PatchEmbed(4, 64, 2)
    x = torch.randn(2, 4, 8, 12)
    pe.set_hw(x)
    assert pe.hw == (4, 6)
    assert torch.equal(pe.patchify(x), OH.patchify(x, 2))
    assert torch.equal(pe.unpatchify(pe.patchify(x)), x)


@pytest.mark.parametrize("steps,shift", [(25, 1.0), (25, 3.0), (10, 1.0), (64, 2.5)])
def test_scheduler_matches_golden(golden_dir, steps, shift):
    d = np.load(os.path.join(golden_dir, "scheduler.npz"))
    s = nb.FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=shift)
    assert s.config.num_train_timesteps == 1000 and s.shift == shift
    s.set_timesteps(steps)
    assert np.array_equal(np.asarray(s.timesteps), d[f"t_{steps}_{shift}"])
    assert np.array_equal(np.asarray(s.sigmas, dtype=np.float64), d[f"s_{steps}_{shift}"])
    ts, sig = OS.schedule(steps, shift=shift)
    assert np.array_equal(np.asarray(s.timesteps), ts) and s.sigmas == sig
    assert s.index_for_timestep(s.timesteps[min(3, steps - 1)]) == min(3, steps - 1)
    s2 = nb.FlowMatchEulerDiscreteScheduler(shift=1.0)
    s2.set_shift(shift)
    s2.set_timesteps(steps)
    if shift == 1.0:
        assert s2.sigmas == s.sigmas


def test_partition_matches_oracle():
    for n in (1024, 2048, 100):
        assert nb.partition.cosine_num_preds(n) == OP.cosine_num_preds(n)
        assert nb.partition.equal_subset_sizes(n) == OP.equal_subset_sizes(n)
    order = torch.stack([torch.randperm(64) for _ in range(3)])
    sets = nb.partition.split_order(order, nb.partition.cosine_num_preds(64, 8))
    assert sum(s.shape[1] for s in sets) == 64 and all(s.shape[0] == 3 and s.shape[2] == 1 for s in sets)
    assert torch.equal(torch.cat(sets, dim=1)[..., 0], order)


def test_guidance_scaler_surface():
    g = nb.GuidanceScaler(guidance_scale=5.0, min_guidance_scale=2.0, guidance_trunc=100)
    g.decay_guidance_scale(0.5)
    assert g.guidance_scale == 3.5
    c = g.clone()
    assert c.guidance_scale == 3.5 and c.guidance_trunc == 100
    x = torch.arange(6.0).reshape(2, 3)
    assert torch.equal(g.expand(x), torch.cat([x, x]))
    assert nb.GuidanceScaler().expand(x) is x
    g3 = nb.GuidanceScaler(guidance_scale=2, image_guidance_scale=1.5)  # third pass (guidance_scaler.py:32-35,46-50)
    assert g3.extra_pass and torch.equal(g3.expand(x), torch.cat([x, x, x])) and g3.clone().image_guidance_scale == 1.5
    assert not g.extra_pass and nb.GuidanceScaler(guidance_scale=2, spatiotemporal_guidance_scale=0.5).extra_pass


def test_shard_range_covers_everything():
    for total, world in [(64, 8), (10, 4), (3, 8), (256, 8), (1, 1)]:
        spans = [nb.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _gather_worker(rank, world, total, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = nb.shard_range(total, rank, world)
    local = torch.arange(total, dtype=torch.float32)[lo:hi].reshape(-1, 1, 1).expand(-1, 4, 3).contiguous()
    out = nb.gather_shards(local, total)
    q.put((rank, out[:, 0, 0].tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 5])
def test_gather_shards_world_size_2_gloo(total):
    """The N>1 path's only collective (one all-gather of the generated points), ragged shards included."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + total) % 400
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, total, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, vals in res:
        assert vals == [float(i) for i in range(total)]


def test_standard_point_cloud_generation_contract():
    """Post-processing of pipeline_nova_pointcloud_gen.py:271-294: subset / tile to num_points, tanh, noise, clamp."""
    import nova_pointcloud_b200 as nb

    pc = torch.randn(50, 3) * 2
    up = nb.standard_point_cloud_generation(pc, 120, noise_scale=0.0)
    assert up.shape == (120, 3) and torch.equal(up, torch.tanh(pc.repeat(3, 1)[:120]))  # 120 // 50 + 1 copies, cut
    g = torch.Generator().manual_seed(3)
    down = nb.standard_point_cloud_generation(pc, 20, generator=g, noise_scale=0.0)
    keep = torch.randperm(50, generator=torch.Generator().manual_seed(3))[:20]
    assert torch.equal(down, torch.tanh(pc[keep]))
    same = nb.standard_point_cloud_generation(pc, 50, generator=torch.Generator().manual_seed(4))
    assert same.shape == (50, 3) and float(same.abs().max()) <= 1.0
    assert 0.05 < float((same - torch.tanh(pc)).std()) < 0.2  # 0.1-sigma noise, clipped at the box


def test_module_copies_and_pickles_without_its_handle():
    """copy.deepcopy (the reference's ModelEMA), pickle and torch.save of a head that has been used: the packed-weights
    handle (a ctypes pointer into device memory) is never copied; load_state_dict / .to() / invalidate() drop it."""
    import copy
    import ctypes
    import io
    import pickle

    import nova_pointcloud_b200 as nb

    class FakeHandle:  # what a used head holds; a bare c_void_p cannot be pickled or deep-copied
        def __init__(self):
            self.ptr, self.closed = ctypes.c_void_p(1234), False

        def close(self):
            self.closed = True

    head = nb.DiffusionMLP(1, 256, 256, patch_size=1, image_dim=3)
    head._handle, head._handle_key = FakeHandle(), ("key",)
    clone = copy.deepcopy(head)
    assert clone._handle is None and clone._handle_key is None and head._handle is not None
    assert all(torch.equal(a, b) for a, b in zip(clone.state_dict().values(), head.state_dict().values()))
    pickle.loads(pickle.dumps(head))
    torch.save(head, io.BytesIO())
    fake = head._handle
    head.load_state_dict(clone.state_dict())
    assert head._handle is None and fake.closed  # stale packed weights can never be served after a load
    head._handle, head._handle_key = FakeHandle(), ("key",)
    head.to(torch.bfloat16)
    assert head._handle is None
    head._handle, head._handle_key = FakeHandle(), ("key",)
    head.invalidate()  # explicit hook for in-place updates through .data (EMA code), which no version counter sees
    assert head._handle is None
