"""CPU suite, part 3: the fold-sharded sweep at world_size 2 over gloo (no GPU, no data-path collective)."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
from mr_gan_b200 import sweep
dist.init_process_group("gloo")
rank = dist.get_rank()
# one seed for the whole sweep: an omitted --seed is drawn on rank 0 and shared (each rank drawing its own would build a
# different dataset and different splits per rank), an explicit one passes through
s = sweep.shared_seed(None)
got = [None, None]
dist.all_gather_object(got, s)
assert got[0] == got[1] and 0 <= s < 2 ** 31, got
assert sweep.shared_seed(1234 + rank * 0) == 1234
jobs = [dict(D=d, n=6000 if i < 9 else 7100) for i, d in enumerate([400, 800, 1200] * 4)]
seen = []
def train_group(js, dev):
    seen.extend(j['D'] for j in js)
    return [(j['D'] * 10 + rank) for j in js]
res = sweep.run_sharded(jobs, train_group, group_size=2, key=lambda j: j['n'], cost=lambda j: j['D'])
assert [r // 10 for r in res] == [j['D'] for j in jobs], res          # reference loop order kept
owners = [r %% 10 for r in res]
assert set(owners) == {0, 1}, owners                                   # both ranks did work
assert len(seen) == owners.count(rank)                                # each fold trained exactly once
load = [sum(j['D'] for j, o in zip(jobs, owners) if o == r) for r in (0, 1)]
assert abs(load[0] - load[1]) <= 2400, load                           # longest-first balance
dist.barrier(); dist.destroy_process_group()
print("OK", rank)
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_fold_sharding_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "OK %d" % r in o
