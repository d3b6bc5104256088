"""torchrun --nproc-per-node N scripts/ddp_check.py : on real GPUs over NCCL, averaged per-rank
gradients of the B200 TopologicalGNN == gradients of the concatenated batch computed by one rank,
and the replicas stay bit-identical after optimizer steps (eager and CUDA-graphed)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, ".")
from gnn_qot_estimation_b200 import TopologicalGNN, synthetic
from gnn_qot_estimation_b200.distributed import GraphDataParallel, shard_range
from gnn_qot_estimation_b200.graphed import GraphedTrainStep

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
backend = os.environ.get("QOT_DDP_BACKEND", "nccl")     # "gloo": N ranks sharing cuda:0 (1-GPU debugging)
if backend == "gloo":
    local = 0
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group(backend, **({"device_id": dev} if backend == "nccl" else {}))
def mark(msg):
    print(f"[rank {rank}] {msg}", flush=True)
G = 64 * world
store = synthetic.nsfnet_store(G, seed=0).to(dev)
torch.manual_seed(100 + rank)                          # different init per rank: broadcast must fix it
model = TopologicalGNN(14, 16, 3, edge_dim=4, dropout_p=0.0).to(dev)
ddp = GraphDataParallel(model)
mark('broadcast done')
crit = torch.nn.SmoothL1Loss()
r = shard_range(G, rank, world)
b = store.collate(r)
ddp.zero_grad(); crit(ddp(b), b.y.view(-1, 3)).backward(); ddp.sync_gradients()
got = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
mark('eager ddp step done')
ref_model = TopologicalGNN(14, 16, 3, edge_dim=4, dropout_p=0.0).to(dev)
ref_model.load_state_dict(model.state_dict())
fb = store.collate(range(0, G))
crit(ref_model(fb), fb.y.view(-1, 3)).backward()
ref = torch.cat([p.grad.reshape(-1) for p in ref_model.parameters()])
err = float((got - ref).abs().max() / ref.abs().max())
opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
opt.step()
g = GraphedTrainStep(model, opt, crit, b, ddp=ddp, warmup=2)
mark(f'graphs captured exchange={g.exchange}')
for _ in range(5):
    loss = g.step(b)
mark('graphed steps done')
flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
ref0 = flat.clone(); dist.broadcast(ref0, src=0)
same = bool(torch.equal(flat, ref0))
res = torch.tensor([err, 0.0 if same else 1.0], device=dev, dtype=torch.float64)
dist.all_reduce(res, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"exchange={g.exchange}")
    print(f"ddp_check world={world}: grad rel err vs concatenated batch {float(res[0]):.3e} (bar 1e-5); "
          f"replicas identical after graphed steps: {float(res[1]) == 0.0}; loss {float(loss):.6f}")
    assert float(res[0]) <= 1e-5 and float(res[1]) == 0.0
# a trainer WITHOUT GraphDataParallel on one rank of the job (bench.py's rank-0-only blocks): its fused update must not
# enter a collective -- the other rank is not there (it waits at the barrier below); this used to hang in the rendezvous
if rank == 0:
    from gnn_qot_estimation_b200.graphed import GraphedStepCache
    solo = TopologicalGNN(14, 16, 3, edge_dim=4, dropout_p=0.0).to(dev).train()
    sopt = torch.optim.SGD(solo.parameters(), lr=0.1, momentum=0.9)
    cache = GraphedStepCache(solo, sopt, lambda m, bb: crit(m(bb), bb.y.view(-1, 3)))
    for _ in range(3):
        cache.step(b)
    torch.cuda.synchronize()
    print("solo trainer on rank 0 stepped without a collective: True")
dist.barrier()
dist.destroy_process_group()
