"""CPU, world_size 2, gloo: the N > 1 host logic of the data-parallel training step — every rank flattens the same
parameter set into the same layout (FlatState), ONE all-reduce (SUM) of the flat gradient buffer followed by the
1/world scale inside AdamW reproduces the single-process gradient of the concatenated batch (what nn.DataParallel's
reduce-add gives the reference, train_CNN.py:185-186), and bench.py's clip sharding gives every rank its own clips."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, oracle, pkg


def _worker(rank: int, world: int, port: int, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        m = pkg()
        train = __import__("importlib").import_module("2023-tifs-istvt_b200.train")
        torch.manual_seed(0)
        vit = m.DSTTr(19, 1, 1, 2, depth=1)              # small on-path parameter set, same on every rank
        model = torch.nn.Module()
        model.vit = vit
        state = train.FlatState(model, train_entry_flow=False)
        names = list(state.names)
        # every parameter is now a view of the flat buffer
        base = state.params.data_ptr()
        assert all(base <= p.data_ptr() < base + state.params.numel() * 4 for p in vit.parameters())
        # rank-local "gradient": deterministic function of (rank, position)
        n = state.grads.numel()
        state.grads.copy_(torch.arange(n, dtype=torch.float32) * 1e-6 + (rank + 1))
        dist.all_reduce(state.grads)                      # SUM, as Trainer.step does
        want_sum = torch.arange(n, dtype=torch.float32) * 1e-6 * world + sum(r + 1 for r in range(world))
        assert torch.allclose(state.grads, want_sum)
        # AdamW with grad_scale = 1/world == AdamW on the mean gradient (oracle restatement, plain torch on CPU)
        O = oracle()
        p_ref = state.params.clone()
        O.adamw_update(p_ref, want_sum / world, torch.zeros(n), torch.zeros(n), 1, 1e-3, weight_decay=0.01)
        p_scaled = state.params.clone()
        O.adamw_update(p_scaled, state.grads * (1.0 / world), torch.zeros(n), torch.zeros(n), 1, 1e-3, weight_decay=0.01)
        assert torch.allclose(p_ref, p_scaled, rtol=0, atol=1e-5)
        # Trainer.step's two gradient buckets (transformer slots reduced asynchronously while the entry-flow backward
        # runs, entry-flow slots at the end) == one all-reduce of the whole flat buffer
        x = torch.nn.Module()
        for nm in ("conv1", "bn1", "conv2", "bn2", "block1", "block2", "block3"):
            setattr(x, nm, torch.nn.Linear(3, 5))
        model.xcep = torch.nn.Module()
        model.xcep.model = x
        st2 = train.FlatState(model, train_entry_flow=True)
        assert st2.vit_offset == 7 * (16 + 8) and st2.names[0].startswith("xcep.") and st2.names[-1].startswith("vit.")
        n2 = st2.grads.numel()
        st2.grads.copy_(torch.arange(n2, dtype=torch.float32) * 1e-6 - rank)
        whole = st2.grads.clone()
        dist.all_reduce(whole)
        work = dist.all_reduce(st2.grads[st2.vit_offset:], async_op=True)
        dist.all_reduce(st2.grads[:st2.vit_offset])
        work.wait()
        assert torch.equal(st2.grads, whole)
        # layouts agree across ranks
        gathered = [None] * world
        dist.all_gather_object(gathered, (names, n))
        assert all(g == gathered[0] for g in gathered)
        # bench.py shards clips by rank: different seeds -> different clips, no overlap to reduce
        g0 = torch.rand(4, generator=torch.Generator().manual_seed(1234 + rank))
        clips = [None] * world
        dist.all_gather_object(clips, g0.tolist())
        assert clips[0] != clips[1]
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, "ok"), (1, "ok")], res
