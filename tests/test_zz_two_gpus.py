"""What needs two GPUs on one box (skipped otherwise): strand_shift over contig-sharded ranks and tags_in_regions over sample-sharded
ranks as separate PROCESSES over NCCL (torchrun, the way bench.py --gpus N runs), each checked against one GPU bit for bit by the
tool itself (src/strand_shift.cpp:196-259, src/tags_in_regions.cpp:131-199)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        from unipeak_b200.capi import lib
        return lib().upb_device_count()
    except Exception:
        return 0


def _torchrun(script, args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", script)] + args
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    assert p.returncode == 0, p.stderr.decode()[-3000:]
    lines = [l for l in p.stdout.decode().splitlines() if l.startswith("{")]
    assert lines, p.stdout.decode()[-2000:]
    return json.loads(lines[-1])


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_strand_shift_sharded_over_nccl_equals_one_gpu():
    out = _torchrun("shift_sharded_bench.py", ["--tags", "4000000", "--reps", "1"], 29611)
    assert out["n_gpus"] == 2 and out["identical_to_one_gpu"] and out["regions_tested"] > 100


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_tags_in_regions_sharded_over_nccl_equals_one_gpu():
    out = _torchrun("tir_sharded_bench.py", ["--samples", "6", "--tags", "3000000", "--reps", "1"], 29612)
    assert out["n_gpus"] == 2 and out["columns_identical_to_one_gpu"] and out["regions"] > 100 and out["tags_counted"] > 0
