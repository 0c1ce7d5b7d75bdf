"""Multi-GPU parity (needs >= 2 visible GPUs, skipped otherwise): partitioned inference and partitioned
training of one scene against the single-GPU result, each as a 2-rank NCCL job."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, port, *args):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script)] + [str(a) for a in args]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_partitioned_inference_matches_single_gpu():
    out = _torchrun("check_partition.py", 29521, 8000)
    assert "PARTITIONED_INFERENCE world=2" in out


def test_partitioned_training_step_matches_single_gpu():
    out = _torchrun("check_partition_train.py", 29522, 8000)
    assert "PARTITIONED_TRAINING world=2" in out


def test_partitioned_updated_training_step_matches_single_gpu():
    out = _torchrun("check_partition_upd.py", 29523, 6000)
    assert "PARTITIONED_UPDATED_TRAINING world=2" in out


def test_sharded_scene_build_matches_single_gpu():
    """Sharded build (no rank holds the whole graph), boundary-first order, overlapped exchange: inference and one
    training step on a Delaunay scene and on the benchmark's lattice scene."""
    out = _torchrun("check_partition_scene.py", 29524, 8000)
    assert "PARTITIONED_SCENE world=2 ok" in out


def test_sharded_scene_single_rank():
    """The same sharded code path with one shard (runs on a 1-GPU box too)."""
    cmd = [sys.executable, os.path.join(ROOT, "tools", "check_partition_scene.py"), "3000"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "PARTITIONED_SCENE world=1 ok" in r.stdout
