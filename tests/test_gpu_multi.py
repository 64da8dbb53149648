"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the sharded path —
one rank per GPU over NCCL (ShardedApproxCounter) and the single-process `--gpus N` binary —
must return exactly the single-GPU counts."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "approx_counter_b200", "csrc", "approx_counter")

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["APC_ROOT"])
import numpy as np, torch, torch.distributed as dist
from approx_counter_b200 import ShardedApproxCounter, ApproxCounter, host
from oracle import orc
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
n, sl, k = 20001, 100, 16
sample = host.synth_ends(77, 0, n, sl, True)             # same whole sample on every rank
thr = host.adjust_threshold(1.0, 16, k)
with ApproxCounter(torch.cuda.current_device()) as c:     # queries from the whole sample
    c.upload_sample(sample)
    km, ct, nd, hn = c.count_kmers_topn(k, thr, 300)
    whole = c.errorCount(km, k)
s = ShardedApproxCounter()
lo, hi = s.upload_sample(sample)
got = s.errorCount(km, k)
s.close()
ok = np.array_equal(got, whole)
codes, offs = orc.encode_matrix(sample[:1500])
ok = ok and lo % 32 == 0 and (hi - lo) >= n // world - 32 * world
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED_OK" if int(flag.item()) == 1 else "SHARDED_MISMATCH", int(got[0]), int(whole[0]))
dist.destroy_process_group()
'''


def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_counter_nccl(built, tmp_path, world):
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, APC_ROOT=ROOT)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-3000:]
    assert "SHARDED_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


@pytest.mark.parametrize("gpus", [2, 3, 8])
def test_binary_multi_gpu_matches_single(built, tmp_path, gpus):
    if n_gpus() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    from approx_counter_b200 import host
    path = tmp_path / "reads.fa"
    host.synth_write(path, 31, 5003, 100)
    outs = {}
    # --ingest device (the default): GPU 0 parses and samples, the others fetch their shard of the sample from it
    # (apc_upload_sample_peer); --ingest host: every shard is uploaded from the host's copy of the sample
    for g, ingest in ((1, "device"), (gpus, "device"), (gpus, "host")):
        o = tmp_path / f"o{g}{ingest}"
        p = subprocess.run([BIN, "-k", "16", "-sn", "5003", "-sl", "100", "-lim", "200", "-e", str(tmp_path / f"e{g}{ingest}"),
                            "-o", str(o), "--gpus", str(g), "--ingest", ingest, "-v", "2", str(path)],
                           capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr
        assert ("indexed on the GPU" in p.stdout) == (ingest == "device")
        outs[g, ingest] = [(tmp_path / f"{pre}{g}{ingest}_0.{end}").read_bytes() for pre in ("o", "e") for end in ("start", "end")]
    assert outs[1, "device"] == outs[gpus, "device"] == outs[gpus, "host"]
    assert len(outs[1, "device"][0].splitlines()) == 200
