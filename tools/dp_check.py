"""Data-parallel parity check (run under torchrun on W GPUs):
W ranks x local batch B/W  ==  one GPU at batch B  (same weights, same global batch, same noise stream).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py [--precision tf32]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from mr_gan_b200.engine import FoldGroup
from mr_gan_b200.model import init_disc, init_gen

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp32")
ap.add_argument("--D", type=int, default=300)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
D, Bg = a.D, a.batch
Bl = Bg // world
rng = np.random.default_rng(0)                      # identical on every rank
pD, pG = init_disc(D, rng), init_gen(D, rng)
for p in pD + pG:
    if p.ndim == 1:
        p += (0.1 * rng.standard_normal(p.shape)).astype(np.float32)
steps = [dict(x_lab=rng.standard_normal((Bg, D)).astype(np.float32), labels=rng.integers(0, 6, Bg).astype(np.int32),
              x_unl=rng.standard_normal((Bg, D)).astype(np.float32), z_d=rng.standard_normal((Bg, 100)).astype(np.float32),
              x_unl2=rng.standard_normal((Bg, D)).astype(np.float32), z_g=rng.standard_normal((Bg, 100)).astype(np.float32))
         for _ in range(a.steps)]
uid = [FoldGroup.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
sl = slice(rank * Bl, (rank + 1) * Bl)

def run(batch, dp):
    out = []
    with FoldGroup([(D, max(2 * batch, 64), 16, 12345)], precision=a.precision, batch=batch, device=local) as fg:
        if dp:
            def allgather(b):
                out = [None] * world
                dist.all_gather_object(out, b)
                return out
            fg.dp_init(rank, world, uid[0], allgather=None if os.environ.get("MRGAN_DP_FUSED") == "0" else allgather)
        fg.set_params(0, 0, pD); fg.set_params(0, 1, pG)
        for s in steps:
            sel = sl if dp else slice(None)
            ll, lu, te = fg.train_batch_disc(0, s['x_lab'][sel], s['labels'][sel], s['x_unl'][sel], s['z_d'][sel])
            lg = fg.train_batch_gen(0, s['x_unl2'][sel], s['z_g'][sel])
            out.append((ll, lu, te, lg))
        return np.array(out), fg.get_params(0, 0), fg.get_params(0, 1)

dp_loss, dp_pD, dp_pG = run(Bl, True)
ok = True
if rank == 0:
    ref_loss, ref_pD, ref_pG = run(Bg, False)
    print("losses DP  :", dp_loss.tolist())
    print("losses 1GPU:", ref_loss.tolist())
    rel = np.abs(dp_loss - ref_loss) / (np.abs(ref_loss) + 1e-12)
    print("max rel loss diff %.3e" % rel.max())
    dmax = max(np.abs(x - y).max() for x, y in zip(dp_pD + dp_pG, ref_pD + ref_pG))
    umax = max(np.abs(x - y).max() for x, y in zip(ref_pD + ref_pG, pD + pG))
    print("max |param diff| %.3e (largest update %.3e)" % (dmax, umax))
    tol = 1e-4 if a.precision == "fp32" else 5e-3
    ok = rel.max() < tol
    print("DP PARITY", "OK" if ok else "FAILED", "(world %d, global batch %d, %s)" % (world, Bg, a.precision))
# all ranks hold identical replicas
flat = torch.tensor(np.concatenate([p.ravel() for p in dp_pD + dp_pG]), device="cuda")
ref = flat.clone(); dist.broadcast(ref, src=0)
same = bool((flat == ref).all())
print("rank %d replicas identical to rank 0: %s" % (rank, same))
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if (ok and same) else 1)
