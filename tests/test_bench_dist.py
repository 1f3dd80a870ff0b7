"""N>1 host-side logic of bench.py on CPU (gloo, world_size 2): rendezvous on 127.0.0.1, barrier,
max-over-ranks reduction, rank-0-only output of the reference arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py")] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_dist_selftest_gloo():
    r = _torchrun(["--dist-selftest"], 29541)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                      # rank 0 only
    d = json.loads(lines[0])
    assert d["world"] == 2 and d["max_ms"] == 20.0 and d["sum_streams"] == 2 * 3
    assert d["stream_bases"] == [0, 1000]       # disjoint synthetic camera sets per rank


def test_reference_arm_under_torchrun_prints_once():
    r = _torchrun(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                   "--cpu-seconds", "1", "--config", "C2"], 29542)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
