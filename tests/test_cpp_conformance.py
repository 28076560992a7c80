"""The C++ host layer (include/ismcts: SOSolver / MOSolver with CudaRootParallel / CudaTreeParallel).

tests/cpp/solvertest.cpp restates the reference's solvertest.cpp / gametest.cpp assertions for the CUDA
policies.  Without a GPU only its host part runs (API shape, game classes, loud failure of a search);
with a GPU the whole program runs."""
import os
import subprocess

import pytest

CPP_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp")


@pytest.fixture(scope="module")
def solvertest(engine_lib):
    subprocess.run(["make", "-C", CPP_DIR], check=True, capture_output=True)
    return os.path.join(CPP_DIR, "solvertest")


def test_host_layer_compiles_and_host_checks_pass(solvertest):
    out = subprocess.run([solvertest, "--host"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failures" in out.stdout


@pytest.mark.gpu
def test_solver_conformance_on_device(solvertest):
    out = subprocess.run([solvertest], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failures" in out.stdout


@pytest.mark.gpu
def test_set_devices_shards_trees_over_two_gpus(solvertest):
    """CudaExecutionPolicy::setDevices({0, 1}): merged statistics equal the one-GPU ones at a fixed seed."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    out = subprocess.run([solvertest, "--devices", "2"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "devices: 2" in out.stdout and "0 failures" in out.stdout
