"""CPU, gloo, world_size 2: the N>1 host logic — sharding by graph and the flat gradient bucket all-reduce
(SURVEY.md §8e).  The model on each rank is the CPU oracle (the CUDA path has no CPU mode); what is tested is the
plumbing that bench.py / engine.TrainStep use on the GPUs."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from glam_b200.engine import FlatGrads
    from glam_b200.synth import make_molecule_batch, shard_by_graph
    from oracle import glam_oracle as O
    torch.manual_seed(0)                                   # identical replicas
    kw = dict(hid_dim_alpha=2, e_dim=16, out_dim=1, message_steps=2, mol_readout="Set2Set", graph_act="CELU")
    model = O.ArchitectureGP(9, 3, **kw)
    full = make_molecule_batch(8, seed=5)
    mine = shard_by_graph(full, rank, world)
    grads = FlatGrads(model.parameters())
    grads.zero()
    loss = torch.nn.functional.mse_loss(model(mine), mine.y)
    loss.backward()
    grads.all_reduce_mean(world)
    # the same step with the gathered bucket split in two: the parameters behind the message stack are reduced from a gradient
    # hook while backward is still running (engine.TrainStep at world_size > 1), the rest at the end
    flat_one = grads.flat.clone()
    reduced = [p.grad.clone() for p in model.parameters()]
    g2 = FlatGrads(model.parameters(), gather=True)
    assert g2.enable_early_bucket(model, (O.MessageBlock,))
    early = sum(p.numel() for p in g2.params[g2._early_at:])

    def fwd_bwd():
        g2.zero()
        torch.nn.functional.mse_loss(model(mine), mine.y).backward()
        g2.collect()
        g2.all_reduce_mean(world)
    ok_order = g2.check_early_order(fwd_bwd)
    fired = []
    orig = dist.all_reduce
    dist.all_reduce = lambda t, *a, **k: (fired.append((t.numel(), bool(k.get("async_op")))), orig(t, *a, **k))[1]
    try:
        fwd_bwd()
    finally:
        dist.all_reduce = orig
    err2 = (g2.flat - flat_one).abs().max().item()
    g2.disable_early_bucket()
    if rank == 0:
        ref = O.ArchitectureGP(9, 3, **kw)
        ref.load_state_dict(model.state_dict())
        torch.nn.functional.mse_loss(ref(full), full.y).backward()
        err = max((g - q.grad).abs().max().item() for g, q in zip(reduced, ref.parameters()))
        torch.save({"err": err, "nodes": mine.num_nodes, "flat": grads.flat.numel(), "err2": err2, "ok_order": ok_order, "fired": fired,
                    "early": early}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_step_matches_full_batch(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["err"] < 1e-6, res
    assert res["flat"] > 1000
    # two-bucket path: same reduced gradients; the early bucket (everything behind the message stack) went asynchronously from the
    # hook BEFORE the rest, and its gradients were final before the stack's
    assert res["ok_order"] and res["err2"] < 1e-7, res
    assert res["fired"] == [(res["early"], True), (res["flat"] - res["early"], False)], res
    assert res["early"] > res["flat"] // 2


def test_shards_partition_the_batch():
    sys.path.insert(0, ROOT)
    from glam_b200.synth import make_molecule_batch, shard_by_graph
    b = make_molecule_batch(37, seed=2)
    for world in (2, 4, 8):
        parts = [shard_by_graph(b, r, world) for r in range(world)]
        assert sum(p.num_graphs for p in parts) == 37
        assert torch.equal(torch.cat([p.x for p in parts]), b.x)
        assert torch.equal(torch.cat([p.y for p in parts]), b.y)
        assert sum(p.num_edges for p in parts) == b.num_edges
