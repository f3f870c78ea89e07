"""rbis_batch_stats_allreduce: the C ABI's statistics reduction for a sharded ensemble over a caller-supplied ncclComm_t
(SURVEY.md 8b export list, 8e protocol).  tests/cpp/stats_allreduce.cpp is the C++ host: one NCCL rank per visible GPU
(two when the box has them, one otherwise -- same code path), totals bit-identical on every rank and to the single-GPU
reduction.  The Python wrapper is checked the same way through pronto_b200.nccl.Communicator."""
import os
import subprocess

import numpy as np
import pytest

from common import gpu_streams, nominal_q, scenario

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
EXE = os.path.join(BUILD, "stats_allreduce")


@pytest.fixture(scope="module")
def exe(rbis_lib):
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "stats_allreduce.cpp")
    libdir = os.path.join(ROOT, "pronto_b200", "lib")
    deps = [src, os.path.join(ROOT, "include", "rbis_batch.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-Wextra", "-I/usr/local/cuda/include", "-o", EXE, src, f"-L{libdir}",
                               "-lrbis_b200", f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64", "-lcudart", "-lnccl", "-lpthread"])
    return EXE


def test_cpp_host_compiles_and_fails_loudly_without_a_gpu(exe):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 5 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_cpp_host_allreduce_is_bit_identical_to_single_gpu(exe):
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "bit-identical" in r.stdout


@pytest.mark.gpu
def test_python_wrapper_matches_stats_plus_reduce_chunks():
    from pronto_b200 import RBISBatch, reduce_chunks, synth
    from pronto_b200.nccl import Communicator

    N, T, CH = 3000, 20, 256
    sc = scenario(N, T)
    st = sc["st"]
    tv, tq = synth.truth_state_at(sc["truth"], T - 1)
    with RBISBatch(N) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(sc["vec"], sc["quat"], sc["cov"])
        b.run_fused(st["events"], imu=st["imu"], streams=gpu_streams(st))
        chunks, _ = b.stats(tv, tq, chunk=CH)
        ref = reduce_chunks(chunks)
        tot0, tab0 = b.stats_allreduce(None, tv, tq, 0, chunks.shape[0], chunk=CH, want_table=True)
        comm = Communicator(0, 1, 0)  # a one-rank communicator: the NCCL path on a single GPU
        tot1, tab1 = b.stats_allreduce(comm, tv, tq, 0, chunks.shape[0], chunk=CH, want_table=True)
        # a shard in the middle of a larger table: its rows land at first_chunk, the rest stays zero
        tot2, tab2 = b.stats_allreduce(comm, tv, tq, 5, chunks.shape[0] + 9, chunk=CH, want_table=True)
        comm.close()
    assert np.array_equal(tot0, ref) and np.array_equal(tab0, chunks)
    assert np.array_equal(tot1, ref) and np.array_equal(tab1, chunks)
    assert np.array_equal(tab2[5:5 + chunks.shape[0]], chunks) and not tab2[:5].any() and not tab2[5 + chunks.shape[0]:].any()
    assert np.array_equal(tot2, ref)
