"""Multi-GPU tests (need at least two B200s: run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`;
skipped on a single-GPU box).  The C program tests/c/sharded_ranks.c drives the C multi-GPU driver of
libliquid_cuda.so the way a C host program would: one process per GPU, the NCCL unique id through a file."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _n_gpus():
    sys.path.insert(0, ROOT)
    import liquid_dsp_b200 as L
    return L.device_count()


def _build(tmp_path):
    exe = tmp_path / "sharded_ranks"
    libdir = os.path.join(ROOT, "liquid-dsp_b200")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "sharded_ranks.c"), "-o", str(exe), "-L", libdir,
                           "-lliquid_cuda", "-lm", "-Wl,-rpath," + libdir])
    return exe


@pytest.mark.parametrize("M,m,frames", [(1024, 7, 4096), (4096, 9, 1024), (64, 4, 4000)])
def test_c_program_drives_sharded_analyzer(tmp_path, M, m, frames):
    world = min(_n_gpus(), 8)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    exe = _build(tmp_path)
    idfile = tmp_path / ("id_%d_%d" % (M, m))
    procs = [subprocess.Popen([str(exe), str(r), str(world), str(idfile), str(M), str(m), str(frames)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        print(o, end="")
        assert p.returncode == 0, "rank %d: %s" % (r, o)
        assert "OK" in o
