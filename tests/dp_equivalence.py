"""Data-parallel gradient equivalence on NCCL (run under torchrun, one rank per GPU; SURVEY.md section 4
"Distributed").  Every rank owns a DIFFERENT design; after ``DesignStep.run`` the rank's ``.grad`` must equal the
mean of the single-GPU gradients of all ranks' designs (rtol 1e-5), for the eager bucketed path, for the CUDA-graph
path with the all-reduces captured inside (default) and for the graph + one exchange after the replay
(TM_DP_GRAPH=0).  Usage: torchrun --nproc-per-node 2 tests/dp_equivalence.py"""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "multimodal-fusion-based-pre-routing-timing-prediction-_b200"
for p in (ROOT, os.path.join(ROOT, PKG)):
    sys.path.insert(0, p)
importlib.import_module(PKG)
import tm_engine  # noqa: E402
import tm_synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
cfg = tm_synth.CONFIGS[os.environ.get("TM_DP_TEST_CONFIG", "tiny")]
designs = [tm_synth.make_design(seed=10 + r, **cfg) for r in range(world)]
model, cnn = tm_engine.build_models(designs[0].map_size, seed=0, device=dev)
params = [p for p in list(model.parameters()) + list(cnn.parameters())]


def grads_of(step, batch, replay=None):
    for p in params:
        p.grad = None
    (replay or (lambda: step.run(batch)))()
    torch.cuda.synchronize()
    return [None if p.grad is None else p.grad.detach().clone() for p in params]


# single-GPU gradients of every design, computed locally on this rank: the expectation is their mean
single = tm_engine.DesignStep(model, cnn)
per_design = [grads_of(single, tm_engine.DesignBatch.from_synth(d, dev)) for d in designs]
expect = [None if g[0] is None else torch.stack(g).mean(0) for g in zip(*per_design)]


def check(got, what):
    for i, (a, b) in enumerate(zip(got, expect)):
        assert (a is None) == (b is None), (what, i)
        if a is None:
            continue
        tol = 1e-5 * b.abs() + 1e-6 * b.abs().max()
        bad = int(((a - b).abs() > tol).sum())
        assert bad == 0, f"{what}: parameter {i}: {bad}/{a.numel()} elements differ from the mean of the single-GPU gradients"


mine = tm_engine.DesignBatch.from_synth(designs[rank], dev)
dp = tm_engine.DesignStep(model, cnn, process_group=dist.group.WORLD, world_size=world)
check(grads_of(dp, mine), "eager bucketed all-reduce")
replay = dp.capture(mine)
check(grads_of(dp, mine, replay), "CUDA graph with the all-reduces captured inside")
check(grads_of(dp, mine, replay), "second replay")
os.environ["TM_DP_GRAPH"] = "0"
replay0 = dp.capture(mine)
check(grads_of(dp, mine, replay0), "CUDA graph + exchange after the replay")
del os.environ["TM_DP_GRAPH"]
dp.close()
dist.barrier()
torch.cuda.synchronize()
dist.destroy_process_group()
print(f"dp equivalence ok rank {rank}/{world}", flush=True)
