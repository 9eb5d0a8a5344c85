"""CPU, world_size 2, gloo: the N>1 host logic of bench.py -- rank-sharded synthetic inputs, barrier, and the
max-over-ranks reduction of the timed region.  (The data path itself has no collective: SURVEY.md section 8e.)"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(1234 + rank)                  # bench.py: per-rank input seed
    x = torch.randn(2, 3, 8, 8, generator=g)
    ms = torch.tensor([10.0 + 5.0 * rank])                          # pretend device time of this rank
    dist.barrier()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    chk = torch.tensor([x.sum().item()])
    allc = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(allc, chk)
    q.put((rank, ms.item(), [c.item() for c in allc]))
    dist.destroy_process_group()


def test_two_rank_gloo_max_time_and_distinct_shards():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in ps]
    assert all(r[1] == 15.0 for r in res)                           # max over ranks
    assert res[0][2] == res[1][2] and res[0][2][0] != res[0][2][1]  # ranks hold different images
    B, steps = 32, 20
    value = world * B * steps / (res[0][1] * steps / 1e3)           # whole-job images/s as bench.py computes it
    assert abs(value - world * B / 15e-3) < 1e-6


def _grad_worker(rank, world, port, q):
    """Data-parallel gradient exchange of the KD step: ONE all-reduce over the flat grad arena."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from efficientsam3_b200.stage1.optim import FlatAdamW
    torch.manual_seed(0)                                            # identical replicas
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 5, 3), torch.nn.BatchNorm2d(5), torch.nn.Flatten(), torch.nn.Linear(5 * 36, 7))
    opt = FlatAdamW(model, lr=1e-3)
    g = torch.Generator().manual_seed(100 + rank)                   # different local gradients
    for p in model.parameters():
        p.grad.copy_(torch.randn(p.shape, generator=g))
    local = opt.flat_grad.clone()
    n = opt.all_reduce_grads()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    ok = torch.equal(opt.flat_grad, sum(gathered)) and n == world
    views_ok = all(torch.equal(p.grad.reshape(-1), opt.flat_grad[o:o + p.numel()]) for p, o in zip(opt.params, opt.offsets))
    q.put((rank, bool(ok), bool(views_ok), opt.flat_grad.sum().item()))
    dist.destroy_process_group()


def test_two_rank_flat_gradient_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in ps]
    assert all(r[1] and r[2] for r in res)
    assert res[0][3] == res[1][3]                                   # both ranks hold the same summed arena


def _student_dp_worker(rank, world, port, q):
    """Data-parallel KD iteration of the native student (ops emulated in fp64, frozen BN so that ranks are exchangeable):
    train-mode forward -> KD loss -> native backward INTO the flat arena -> one all-reduce.  The averaged arena must equal
    the gradient of the same model on the concatenated batch."""
    import sys
    from types import SimpleNamespace as NS
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(4)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import emu_ops
    from efficientsam3_b200 import ops
    from efficientsam3_b200.stage1.model import build_image_student_model
    from efficientsam3_b200.stage1.optim import FlatAdamW
    from oracle.kd_loss import kd_loss
    from oracle.weights import fill_state_dict

    class _Patch:
        def setattr(self, obj, name, value):
            setattr(obj, name, value)

    emu_ops.install(_Patch())
    emu_ops.BF = emu_ops.CD = torch.float64
    ops.ACT_DTYPE = torch.float64
    img, embed, b = 160, 12, 1
    cfg = NS(MODEL=NS(BACKBONE="efficientvit_b0"), DATA=NS(IMG_SIZE=img), DISTILL=NS(EMBED_DIM=1024, EMBED_SIZE=embed))

    def make():
        m = build_image_student_model(cfg)
        m.load_state_dict(fill_state_dict(m.state_dict(), 5))
        m.train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
                mod.eval()
        return m

    g = torch.Generator().manual_seed(7)
    x_all = torch.randn(world * b, 3, img, img, generator=g)
    t_all = torch.randn(world * b, 1024, embed, embed, generator=g).double()
    sizes = [(3, img, img)] * (world * b)

    # the global-batch reference first (before this rank takes part in any collective)
    ref = make()
    ropt = FlatAdamW(ref, lr=1e-3)
    ropt.zero_grad()
    rl, _, _ = kd_loss(ref(x_all), t_all, img, sizes, 1.0)
    rl.backward()

    m = make()
    opt = FlatAdamW(m, lr=1e-3)
    opt.zero_grad()
    sl = slice(rank * b, (rank + 1) * b)
    loss, _, _ = kd_loss(m(x_all[sl]), t_all[sl], img, sizes[sl], 1.0)
    opt.begin_backward(True)                           # as kd_train_step does: the head range is exchanged from inside the backward
    loss.backward()                                    # accumulates into opt.flat_grad through the p.grad views
    assert opt.early_exchanges == 1 and len(opt._pending) == 1, (opt.early_exchanges, opt._pending)
    n = opt.all_reduce_grads()                         # the two remaining ranges + wait for the early one
    assert not opt._pending
    mean_grad = opt.flat_grad / n
    err = ((mean_grad - ropt.flat_grad).norm() / ropt.flat_grad.norm()).item()
    q.put((rank, n, err, float(mean_grad.abs().sum())))
    dist.destroy_process_group()


def test_two_rank_student_kd_iteration_matches_the_global_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_student_dp_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=300) for _ in range(world))
    [p.join(timeout=60) for p in ps]
    assert all(r[1] == world for r in res)
    assert all(r[2] < 1e-5 for r in res), res             # fp32 accumulators in the arena are the only rounding left
    assert res[0][3] == res[1][3] and res[0][3] > 0         # both ranks hold the same averaged arena


def _student_ddp_worker(rank, world, port, q):
    """The reference's own arrangement: torch DistributedDataParallel around the (native) student, loss.backward(), gradients
    averaged by DDP's bucketed all-reduce (train_image_encoder_stage1.py:96-101).  The native forward is ONE autograd node fed
    by the parameters, so DDP's per-parameter hooks fire as for any module."""
    import sys
    from types import SimpleNamespace as NS
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(4)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import emu_ops
    from efficientsam3_b200 import ops
    from efficientsam3_b200.stage1.model import build_image_student_model
    from oracle.kd_loss import kd_loss
    from oracle.weights import fill_state_dict

    class _Patch:
        def setattr(self, obj, name, value):
            setattr(obj, name, value)

    emu_ops.install(_Patch())
    emu_ops.BF = emu_ops.CD = torch.float64
    ops.ACT_DTYPE = torch.float64
    img, embed, b = 160, 12, 1
    cfg = NS(MODEL=NS(BACKBONE="efficientvit_b0"), DATA=NS(IMG_SIZE=img), DISTILL=NS(EMBED_DIM=1024, EMBED_SIZE=embed))

    def make():
        m = build_image_student_model(cfg)
        m.load_state_dict(fill_state_dict(m.state_dict(), 5))
        m.train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
                mod.eval()
        return m

    g = torch.Generator().manual_seed(7)
    x_all = torch.randn(world * b, 3, img, img, generator=g)
    t_all = torch.randn(world * b, 1024, embed, embed, generator=g).double()
    sizes = [(3, img, img)] * (world * b)
    ref = make()
    rl, _, _ = kd_loss(ref(x_all), t_all, img, sizes, 1.0)
    rl.backward()

    ddp = torch.nn.parallel.DistributedDataParallel(make(), broadcast_buffers=False)
    sl = slice(rank * b, (rank + 1) * b)
    loss, _, _ = kd_loss(ddp(x_all[sl]), t_all[sl], img, sizes[sl], 1.0)
    loss.backward()
    num = den = 0.0
    for (k, p), (_, r) in zip(ddp.module.named_parameters(), ref.named_parameters()):
        num += (p.grad.double() - r.grad.double()).pow(2).sum().item()
        den += r.grad.double().pow(2).sum().item()
    q.put((rank, (num / den) ** 0.5))
    dist.destroy_process_group()


def test_two_rank_torch_ddp_around_the_native_student():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_student_ddp_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=300) for _ in range(world))
    [p.join(timeout=60) for p in ps]
    assert all(r[1] < 1e-5 for r in res), res


def test_rank_partition_matches_the_reference_sampler():
    """stage1.sharding.shard_indices vs indices recorded from the unmodified MyDistributedSampler (stage1/data/sampler.py)
    for 102 (dataset size, world, rank, epoch, option) combinations incl. padding, drop_last, pair and world > dataset."""
    import json
    from efficientsam3_b200.stage1.sharding import ShardedSampler, shard_indices
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sampler_cases.json")))
    assert len(cases) >= 100
    for c in cases:
        got = shard_indices(c["n"], c["world"], c["rank"], c["epoch"], **c["kw"])
        assert got == c["indices"], c
        s = ShardedSampler(list(range(c["n"])), num_replicas=c["world"], rank=c["rank"], **c["kw"])
        s.set_epoch(c["epoch"])
        assert list(s) == c["indices"] and len(s) == len(c["indices"])
    # every epoch all ranks together cover the dataset, and config 4's shape: 256 images -> 8 ranks x 32
    parts = [shard_indices(256, 8, r, epoch=2, seed=1) for r in range(8)]
    assert all(len(p) == 32 for p in parts) and sorted(sum(parts, [])) == list(range(256))
