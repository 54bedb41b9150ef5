"""CPU tests of bench.py's contract pieces that need no GPU, and of the host-placement helper."""
import json
import os
import subprocess
import sys

from reve_b200 import numa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun (N > 1) rank 0 alone runs the reference arm; the other ranks exit 0 without work or output."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_both_arms_describe_the_same_config():
    """The driver compares the two arms' `config` blocks key by key (round 1: a spurious same_config = false)."""
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    args = argparse.Namespace(workload="720p_x4", tile=200, prepad=10, batch=12)
    a = bench.config_block(args, 1, "x")
    b = bench.config_block(args, 8, "y")
    assert list(a) == list(b) and a["frame"] == [1280, 720] and a["scale"] == 4
    assert "configs[2]" in a["workload"]
    assert bench.metric_name("1080p_x2") == "frames/s animevideov3 x2 1080p->4K"      # BASELINE.json's metric
    # the traffic record the roofline block quotes exists and is self-consistent
    t = bench.measured_traffic(chained=True)
    assert t and t["dram_bytes_per_launch"] == t["dram_bytes_read"] + t["dram_bytes_write"]
    assert t["dram_bytes_per_frame"] > 100 * t["algorithmic_minimum_bytes_per_frame"] / 1.1   # ~110x, stated, not hidden
    json.dumps(a)
    # the same command line gives the same block on both arms, value for value
    args = argparse.Namespace(workload="1080p_x2", tile=200, prepad=10, batch=12, single_process=False, pg="nccl", gpus=1)
    ours = bench.config_block(args, 1, bench.parallelism_string(args, 1))
    ref = bench.config_block(args, args.gpus, bench.parallelism_string(args, args.gpus))
    assert ours == ref and ours["parallelism"] == "segments x1, no collective"


def test_numa_helper_reads_sysfs(tmp_path, monkeypatch):
    """A lane binds itself to the CPUs of the NUMA node its GPU hangs off; unknown topology is a no-op."""
    assert numa._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert numa._parse_cpulist("") == set()
    monkeypatch.setattr(numa, "gpu_pci_bus_id", lambda d: "0000:1b:00.0" if d == 0 else None)
    dev = tmp_path / "bus" / "pci" / "devices" / "0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("1\n")
    node = tmp_path / "devices" / "system" / "node" / "node1"
    node.mkdir(parents=True)
    (node / "cpulist").write_text("16-31\n")
    assert numa.cpus_local_to_gpu(0, sysfs=str(tmp_path)) == set(range(16, 32))
    (dev / "numa_node").write_text("-1\n")                       # single-node hosts
    assert numa.cpus_local_to_gpu(0, sysfs=str(tmp_path)) is None
    assert numa.cpus_local_to_gpu(5, sysfs=str(tmp_path)) is None and numa.bind_thread_to_gpu_node(5) is None
