"""-m gpu: sharded ID-only path == single-GPU engine (world 1 always; world 2 over NCCL when two GPUs are visible)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_sharded_world1_equals_engine():
    r = subprocess.run([sys.executable, os.path.join(HERE, "dist_gpu_check.py")], capture_output=True, text=True, timeout=300)
    assert "DIST_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_sharded_demand_world1_equals_engine():
    """demand-driven step (last layer on the batch's neighbourhood only, row-sparse user gradient) == the dense single-GPU engine"""
    r = subprocess.run([sys.executable, os.path.join(HERE, "dist_gpu_check.py")], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, LLMREC_DIST_DEMAND="1"))
    assert "DIST_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_sharded_demand_world2_equals_engine():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29521", os.path.join(HERE, "dist_gpu_check.py")], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, LLMREC_DIST_DEMAND="1"))
    assert "DIST_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_sharded_world2_equals_engine():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(HERE, "dist_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert "DIST_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
