"""CPU, world_size 2 (gloo): the data-parallel step order of the B200 trainer -- SUM all-reduce of ONE flat gradient
arena (`engine/dp.allreduce_grads`), 1/world folded into the consumer (`dp.pre_scale`), global-norm clip AFTER the
reduction, SGD -- leaves every rank with the parameters torch DistributedDataParallel + the reference's
`clip_gradient` + `optimizer.step()` produce (`pt/engine/trainer.py:92-95,383-386,592-603`), each rank training on
its own images. Both arms run the CPU oracle model (the CUDA model needs a GPU); what is under test is the
reduction / scaling / clipping order that `PTrainer._optimizer_step` implements over the arena."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys, copy, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from torch.nn.parallel import DistributedDataParallel as DDP
from oracle import pt_oracle as O
from probabilisticteacher_b200.engine import dp
torch.set_num_threads(2)
rank = int(os.environ["RANK"])
dist.init_process_group("gloo", rank=rank, world_size=2)
cfg = O.OracleCfg(num_classes=8, base_lr=0.02)
H, W = 64, 96
batch = O.synthetic_batch(1, H, W, 8, 100 + rank, boxes_per_image=3)     # every rank its own image

class S:                                                                   # identical sampling on both arms
    def prio(self, tag, n):
        g = torch.Generator().manual_seed(hash(tag) % 1000 + 7 * rank)
        return torch.rand(n, generator=g)

def loss_of(model):
    losses, _, _, _ = model(batch, branch="supervised")
    return sum(losses.values())

# ---- arm A: torch DDP (gradient averaging) + clip_gradient + SGD, as the reference trains
a = O.OracleRCNN(cfg, seed=5 + rank)            # different initial weights per rank: DDP broadcasts rank 0's
a.sampler = S()
ddp = DDP(a, find_unused_parameters=True)
opt_a = O.make_optimizer(a, cfg)
opt_a.zero_grad()
loss_of(ddp).backward()
O.clip_gradient(a.parameters(), 0.5)            # small clip norm so that the clip is active
opt_a.step()

# ---- arm B: the arena order
b = O.OracleRCNN(cfg, seed=5 + rank)
b.sampler = S()
params = [p for p in b.parameters()]
flat = torch.cat([p.detach().reshape(-1) for p in params])
dp.broadcast_params(flat)                        # PTrainer.__init__: rank 0's arena everywhere
off = 0
with torch.no_grad():
    for p in params:
        p.copy_(flat[off:off + p.numel()].view_as(p)); off += p.numel()
opt_b = O.make_optimizer(b, cfg)
opt_b.zero_grad()
loss_of(b).backward()
train = [p for p in params if p.requires_grad]
g = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in train])
for w in dp.allreduce_grads(g, bucket_elems=1 << 20, async_op=True):      # several buckets, SUM
    w.wait()
g *= dp.pre_scale()                                                         # 1 / world, applied by the consumer
off = 0
for p in train:
    p.grad = g[off:off + p.numel()].view_as(p).clone(); off += p.numel()
O.clip_gradient(train, 0.5)                                                 # clip AFTER the reduction
opt_b.step()

worst = 0.0
for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
    worst = max(worst, float((pa - pb).abs().max() / pa.abs().max().clamp_min(1e-12)))
assert worst < 1e-5, worst
# and the ranks agree with each other (replicas stay identical)
chk = torch.cat([p.detach().reshape(-1)[:64] for p in b.parameters()])
both = [torch.zeros_like(chk) for _ in range(2)]
dist.all_gather(both, chk)
assert torch.equal(both[0], both[1])
dist.destroy_process_group()
print("ok", worst)
'''


def test_arena_allreduce_order_equals_ddp(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29541",
                   PYTHONHASHSEED="0")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out = p.communicate(timeout=600)[0]
        assert p.returncode == 0 and "ok" in out, out[-3000:]
